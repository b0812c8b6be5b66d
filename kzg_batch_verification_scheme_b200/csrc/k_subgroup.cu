// K1b / K1c: the two |x| double-and-add chains of the subgroup check (see k_decompress.cu for the split).
// (An out-of-line build of the Fp products -- one 6 KB copy per kernel so that the doubling loop fits the
// instruction cache -- was measured: 101 ms vs 98 ms inlined at n = 2^20, so the products stay inlined.)
#include "kernels.h"

__device__ __forceinline__ Fp ld_fp(const Fp* p) {
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4 a = s[0], b = s[1], c = s[2];
    Fp r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    r.v[8] = c.x; r.v[9] = c.y; r.v[10] = c.z; r.v[11] = c.w;
    return r;
}
__device__ __forceinline__ void st_fp(Fp* p, const Fp& r) {
    uint4* d = reinterpret_cast<uint4*>(p);
    d[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    d[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
    d[2] = make_uint4(r.v[8], r.v[9], r.v[10], r.v[11]);
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_subgroup_chain1(const Fp* __restrict__ pts, size_t m, Fp* __restrict__ tmp) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    G1Aff p = {ld_fp(pts + 2 * i), ld_fp(pts + 2 * i + 1)};
    if (aff_is_inf(p)) return;                       // infinity or already rejected: nothing to check
    G1Jac t = jac_mul_xabs_aff(p);
    st_fp(tmp + 3 * i, t.X); st_fp(tmp + 3 * i + 1, t.Y); st_fp(tmp + 3 * i + 2, t.Z);
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_subgroup_chain2(Fp* __restrict__ pts, size_t m, const Fp* __restrict__ tmp,
                                                               u8* __restrict__ status, u32* __restrict__ counters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    Fp px = ld_fp(pts + 2 * i);
    {
        Fp py = ld_fp(pts + 2 * i + 1);
        if (fp_is_zero(px) && fp_is_zero(py)) return;
    }
    G1Jac t = {ld_fp(tmp + 3 * i), ld_fp(tmp + 3 * i + 1), ld_fp(tmp + 3 * i + 2)};
    G1Jac q = jac_mul_xabs(t);
    G1Aff p = {px, ld_fp(pts + 2 * i + 1)};
    if (!g1_subgroup_compare(p, q)) {
        Fp z = fp_zero();
        st_fp(pts + 2 * i, z); st_fp(pts + 2 * i + 1, z);
        status[i] = (u8)ST_NOT_IN_G1;
        atomicAdd(counters, 1u);
    }
}


void launch_subgroup_chains(cudaStream_t s, Fp* pts, size_t m, Fp* tmp, uint8_t* status, uint32_t* counters, int mb, int mc) {
    unsigned blocks = (unsigned)((m + 127) / 128);
    if (mb >= 4) k_subgroup_chain1<4><<<blocks, 128, 0, s>>>(pts, m, tmp);
    else if (mb == 3) k_subgroup_chain1<3><<<blocks, 128, 0, s>>>(pts, m, tmp);
    else k_subgroup_chain1<2><<<blocks, 128, 0, s>>>(pts, m, tmp);
    if (mc >= 4) k_subgroup_chain2<4><<<blocks, 128, 0, s>>>(pts, m, tmp, status, counters);
    else if (mc == 3) k_subgroup_chain2<3><<<blocks, 128, 0, s>>>(pts, m, tmp, status, counters);
    else k_subgroup_chain2<2><<<blocks, 128, 0, s>>>(pts, m, tmp, status, counters);
}
