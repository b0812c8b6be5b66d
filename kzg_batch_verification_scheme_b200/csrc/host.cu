// Host orchestration + the C ABI of include/kzgb200.h for the CUDA product library (libkzgb200.so).
// One DeviceSlot per GPU: own stream, workspaces sized for n_max, pinned mailboxes.  A batch is cut into
// contiguous shards (multiples of KZGB_CHUNK proofs) over the slots; only chunk digests (32 B / KZGB_CHUNK proofs),
// the 32-byte root and one 320-byte partial per shard cross the host (BASELINE.json:5: "combined on the
// host, so no NCCL is needed").  There is NO CPU fallback: every arithmetic step runs in a kernel.
#include <cstdio>
#include <chrono>
#include <new>
#include <algorithm>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/kzgb200.h"
#include "kernels.h"
#include "eip4844.cuh"

#define CK(x)                                                                                         \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess) {                                                                      \
            fprintf(stderr, "[kzgb200] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return KZGB_ERROR;                                                                        \
        }                                                                                             \
    } while (0)

// Every entry point leaves the calling thread's current CUDA device as it found it (the library switches devices
// internally; a caller such as PyTorch keeps allocating on "its" device afterwards).
struct ApiDeviceGuard {
    int dev = -1;
    ApiDeviceGuard() { if (cudaGetDevice(&dev) != cudaSuccess) { dev = -1; cudaGetLastError(); } }
    ~ApiDeviceGuard() { if (dev >= 0) cudaSetDevice(dev); }
};
#define KZ_API_GUARD ApiDeviceGuard api_device_guard_

static_assert(KZGB_CHUNK == KZ_FS_CHUNK, "chunk size of the public header and of the device hash must agree");

namespace {

struct SortBuf {
    uint32_t *keys = nullptr, *vals = nullptr, *keys_alt = nullptr, *vals_alt = nullptr, *bucket_start = nullptr;
    uint32_t *count = nullptr, *cursor = nullptr, *tile_sum = nullptr;
    size_t capacity = 0;
};

struct DeviceSlot {
    int device = 0;
    size_t n_max = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;    // high-priority side stream: hashes, challenges and sorts overlap K1
    cudaStream_t stream3 = nullptr, stream4 = nullptr;   // the three sums accumulate/reduce concurrently
    cudaStream_t stream5 = nullptr;    // S1 (stream3: S3, stream4: S2'); all at normal priority -- running S1 and S3 at high
                                       // priority so that the subgroup chains start earlier was slower at every batch size
    // staged inputs (host-pointer API)
    uint8_t *dC = nullptr, *dz = nullptr, *dy = nullptr, *dpi = nullptr;
    Fp* pts = nullptr;                 // 3*n_max + 2 affine points: C | pi | G | phi(pi) | phi(G)
    Fp* k1_tmp = nullptr;              // 3 Fp per point: [|x|]P between the subgroup-check kernels
    uint8_t* status = nullptr;         // 2*n_max
    uint32_t* counters = nullptr;      // [0] bad points [1] bad scalars [2] slice sums outside G1 (batched check)
    uint32_t *leaves = nullptr, *digests = nullptr, *root_words = nullptr;
    uint32_t *r = nullptr, *rz = nullptr, *zs = nullptr, *partials = nullptr, *sum_ry = nullptr;   // zs: GLV halves of rz
    SortBuf sortR, sortZ;
    G1Xyzz *bucketsA = nullptr, *bucketsB = nullptr, *bucketsC = nullptr, *winsums = nullptr;
    size_t max_bucketsR = 0, max_bucketsZ = 0;
    ChunkRecs recs = {nullptr, nullptr, nullptr, nullptr, nullptr};      // sum S2' (and kzgb_g1_msm)
    ChunkRecs recsA = {nullptr, nullptr, nullptr, nullptr, nullptr}, recsB = {nullptr, nullptr, nullptr, nullptr, nullptr};
    G1Xyzz* sg_partial = nullptr;      // bucket reduction scratch of three sums: run sums, totals, slice sums
    size_t sg_cap = 0;                 // entries per sum (the last 256 are the slice sums)
    size_t sg_min = 2;                 // batches of at least this many proofs use the batched subgroup check (0 = never)
    bool sg_batch = false;             // current shard: K1 ran without the per-point chains
    bool head_mode = false;            // current shard: K1 started after the first eighth of C was resident (ev[17])
    G1Jac* sums = nullptr;             // [0] S1 [1] S2' [2] S3 [3] A [4] B
    uint8_t* partial_dev = nullptr;    // 320
    uint8_t* partials_in = nullptr;    // 320 * 64
    uint8_t* scratch = nullptr;        // misc byte buffer (debug ops, artefacts): 1 MiB
    int* result_dev = nullptr;
    G2Lines* lines = nullptr;
    Fp* g1_pt = nullptr;
    int* setup_status = nullptr;
    Fp* comb = nullptr;
    bool comb_built = false;
    // cell batch (config[4]): [tau^j]G1 j<64, lines of G2 and [tau^64]G2, twiddles omega^-t, per-call workspace
    bool cell_ready = false;
    Fp* cell_g1 = nullptr;
    G2Lines* lines_cell = nullptr;
    Fr* cell_W = nullptr;
    uint8_t* d_cells = nullptr;
    uint32_t *d_ci = nullptr, *d_xi = nullptr;
    Fr* cell_coefs = nullptr;
    size_t cell_cap = 0;
    // blob batch: staging for KZ_BLOB_STAGE blobs and their leaf digests
    uint8_t* d_blobs = nullptr;
    uint32_t* blob_leaves = nullptr;
    bool have_ab = false;              // sums[3], sums[4] hold the pairing inputs of the last call
    // Horner-free pairing check (mpair.cuh): line tables of the 2 x 33 fixed G2 multiples, per-call buffers
    G2Aff* mp_chain = nullptr;
    G2Lines *mp_tab = nullptr, *mp_tab_cell = nullptr;
    G1Xyzz *mp_terms = nullptr, *mp_terms_in = nullptr;     // this shard's 66 terms; all shards' terms (root device)
    MpCoef* mp_coef = nullptr;
    Fp12 *mp_part = nullptr, *mp_F = nullptr;
    uint8_t* h_terms = nullptr;        // pinned: 64 x 66 terms
    bool classic = false;              // KZGB_CLASSIC=1: Horner combine + two-pairing kernel (the stage-export path) for every batch
    bool sums_pending = false;         // the last call left slices only: S1, S2', S3 are computed when artefacts are requested
    MsmWorkspace ws_keep[3];           // workspaces of S1, S3, S2' of the last call (for the deferred Horner combine)
    // pinned mailboxes
    uint8_t* h_digests = nullptr;      // 32 * ceil(n_max/KZGB_CHUNK)
    uint32_t* h_small = nullptr;       // 64 words: [0..2] counters, [8..15] root words, [16] result
    uint8_t* h_partial = nullptr;      // 320 * 64
    cudaEvent_t ev[32] = {};
    cudaStream_t stream6 = nullptr;    // high priority: S1 when it starts under K1 of the proofs (s1_early)
    // the fixed tail of every batch -- coefficients, line products, merge, serial check, verdict D2H -- as ONE CUDA graph
    // per (terms buffer, shard count, line table); KZGB_NO_GRAPH=1 launches the four kernels one by one
    struct TailGraph { const void* terms; int n_shards; const void* tab; cudaGraphExec_t exec; };
    std::vector<TailGraph> tail_graphs;
    bool use_graph = true;
    bool s1_early = false;             // current shard: K1 ran commitments first and recorded ev[26] (+ ev[29]) when they were done
    bool c_done_two = false;           // ... on two streams (host pieces alternate): ev[29] as well
    bool s1_early_enabled = true;      // KZGB_S1_EARLY=0 switches the overlap off
    bool s1_early_force = false;       // KZGB_S1_EARLY=3: for every shard of >= 16384 proofs (measurement)
    // current shard (between phase 1 and phase 2)
    const uint8_t *cur_C = nullptr, *cur_z = nullptr, *cur_y = nullptr, *cur_pi = nullptr;
    size_t cur_n = 0;
    bool have_sums = false;
    MsmPlan planR, planZ;
};

}  // namespace

// One persistent host thread per additional device: the per-shard phases of a multi-device batch are issued (and
// waited for) side by side instead of one device after the other.  Job 0 runs on the calling thread.
class SlotPool {
public:
    ~SlotPool() { stop(); }
    void start(size_t n_workers) {
        for (size_t i = 0; i < n_workers; ++i) {
            workers_.emplace_back(new Worker());
            Worker* w = workers_.back().get();
            w->th = std::thread([w] {
                std::unique_lock<std::mutex> lk(w->mu);
                for (;;) {
                    w->cv.wait(lk, [w] { return w->has_job || w->quit; });
                    if (w->quit) return;
                    lk.unlock();
                    w->job();
                    lk.lock();
                    w->has_job = false;
                    w->cv_done.notify_one();
                }
            });
        }
    }
    void stop() {
        for (auto& w : workers_) {
            { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
            w->cv.notify_one();
            if (w->th.joinable()) w->th.join();
        }
        workers_.clear();
    }
    // fn(0) .. fn(n - 1), one job per thread; returns when all are done
    template <class F>
    void run(size_t n, F&& fn) {
        size_t used = n > 0 ? std::min(n - 1, workers_.size()) : 0;
        for (size_t i = 0; i < used; ++i) {
            Worker* w = workers_[i].get();
            { std::lock_guard<std::mutex> lk(w->mu); w->job = [&fn, i] { fn(i + 1); }; w->has_job = true; }
            w->cv.notify_one();
        }
        if (n > 0) fn(0);
        for (size_t g = used + 1; g < n; ++g) fn(g);            // more jobs than threads: the caller finishes them
        for (size_t i = 0; i < used; ++i) {
            Worker* w = workers_[i].get();
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv_done.wait(lk, [w] { return !w->has_job; });
        }
    }
private:
    struct Worker {
        std::thread th;
        std::mutex mu;
        std::condition_variable cv, cv_done;
        std::function<void()> job;
        bool has_job = false, quit = false;
    };
    std::vector<std::unique_ptr<Worker>> workers_;
};

// One lane of the submit / wait pipeline: a complete workspace (DeviceSlot) on device 0 of the context and a host
// thread that drives one batch at a time through it, so the serial tail of batch k (bucket reduction, pairing) runs
// under K1 of batch k+1.
struct Lane {
    DeviceSlot slot;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    enum State { IDLE, RUNNING, DONE } state = IDLE;
    bool quit = false;
    const uint8_t *C = nullptr, *z = nullptr, *y = nullptr, *pi = nullptr;
    size_t n = 0;
    bool on_device = false;
    uint64_t ticket = 0;
    kzgb_ret rc = KZGB_OK;
    bool ok = false;
};

struct kzgb_ctx {
    std::vector<DeviceSlot> slots;
    SlotPool pool;
    std::vector<uint8_t> setup_g1, setup_g2;     // the caller's setup bytes (further workspaces are initialised from them)
    std::vector<std::unique_ptr<Lane>> lanes;
    uint64_t next_ticket = 0;
    std::mutex lane_mu;
    kzgb_artifacts art;
    float msm_ms[4] = {0, 0, 0, 0};
    uint64_t launches_at_create = 0;
    int n_shards_last = 1;
    bool terms_combined = false;       // slot 0 holds the pairing terms of the last batch; A and B are computed on request
    const G1Xyzz* ab_terms = nullptr;  // ... from these terms (device memory of slot 0), ab_shards shards
    int ab_shards = 0;
    bool ab_gather_sum_ry = false;     // sum r_i y_i has to be added up over the shards' devices as well
};

namespace {

template <class T>
cudaError_t dmalloc(T*& p, size_t count) { return cudaMalloc((void**)&p, count * sizeof(T)); }

kzgb_ret slot_alloc_sort(SortBuf& b, size_t cap, size_t buckets) {
    b.capacity = cap;
    CK(dmalloc(b.keys, cap)); CK(dmalloc(b.vals, cap)); CK(dmalloc(b.keys_alt, cap)); CK(dmalloc(b.vals_alt, cap));
    CK(dmalloc(b.bucket_start, buckets + 2)); CK(dmalloc(b.count, buckets + 2)); CK(dmalloc(b.cursor, buckets + 2));
    CK(dmalloc(b.tile_sum, 1024));
    return KZGB_OK;
}

kzgb_ret slot_init(DeviceSlot& s, int device, size_t n_max, const uint8_t* g1m, size_t n1, const uint8_t* g2m, size_t n2) {
    s.device = device;
    s.n_max = n_max;
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    {
        int lo_pri = 0, hi_pri = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
        CK(cudaStreamCreateWithPriority(&s.stream2, cudaStreamNonBlocking, hi_pri));
        CK(cudaStreamCreateWithFlags(&s.stream3, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s.stream4, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&s.stream5, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithPriority(&s.stream6, cudaStreamNonBlocking, hi_pri));
        { const char* e = getenv("KZGB_NO_GRAPH"); s.use_graph = !(e && atoi(e) != 0); }
        { const char* e = getenv("KZGB_S1_EARLY"); s.s1_early_enabled = !(e && atoi(e) == 0); s.s1_early_force = e && atoi(e) == 3; }
    }
    for (auto& e : s.ev) CK(cudaEventCreate(&e));
    CK(dmalloc(s.dC, 48 * (n_max + 1))); CK(dmalloc(s.dz, 32 * (n_max + 1))); CK(dmalloc(s.dy, 32 * (n_max + 1))); CK(dmalloc(s.dpi, 48 * (n_max + 1)));
    CK(dmalloc(s.pts, 2 * (4 * n_max + 8)));            // EIP-4844 mode lays out C | phi(C) | pi | G | phi(pi) | phi(G) for n <= n_max / 2
    CK(dmalloc(s.k1_tmp, 3 * (2 * n_max + 2)));
    CK(dmalloc(s.status, 2 * n_max + 2));
    CK(dmalloc(s.counters, 8));
    size_t nch = (n_max + KZGB_CHUNK - 1) / KZGB_CHUNK;
    CK(dmalloc(s.leaves, 8 * n_max)); CK(dmalloc(s.digests, 8 * nch)); CK(dmalloc(s.root_words, 8));
    CK(dmalloc(s.r, 4 * n_max)); CK(dmalloc(s.rz, 8 * (n_max + 1))); CK(dmalloc(s.zs, 4 * 2 * (n_max + 1)));
    CK(dmalloc(s.partials, 8 * ((n_max + 127) / 128 + 1))); CK(dmalloc(s.sum_ry, 8));
    // the 255-bit sum is GLV-split into 2(n+1) 128-bit scalars.  Window widths depend on n (msm_make_plan),
    // so size every workspace for the worst case over batch sizes up to n_max
    size_t capR = 0, capZ = 0;
    s.max_bucketsR = s.max_bucketsZ = 0;
    for (int k = 1; k <= 64; ++k) {
        size_t nn = n_max * (size_t)k / 64;
        if (nn < 1) nn = 1;
        MsmPlan pr = msm_make_plan(nn, 128), pz = msm_make_plan(2 * (nn + 1), 128);
        capR = std::max(capR, (size_t)pr.W * nn);
        capZ = std::max(capZ, (size_t)pz.W * 2 * (nn + 1));
        s.max_bucketsR = std::max(s.max_bucketsR, (size_t)pr.total_buckets);
        s.max_bucketsZ = std::max(s.max_bucketsZ, (size_t)pz.total_buckets);
    }
    capR += capR / 8 + 4096;
    capZ += capZ / 8 + 8192;
    s.max_bucketsR += 4096; s.max_bucketsZ += 4096;
    if (slot_alloc_sort(s.sortR, capR, s.max_bucketsR + 512)) return KZGB_ERROR;
    if (slot_alloc_sort(s.sortZ, capZ, s.max_bucketsZ + 512)) return KZGB_ERROR;
    CK(dmalloc(s.bucketsA, s.max_bucketsR + 512)); CK(dmalloc(s.bucketsB, s.max_bucketsR + 512));
    {
        // scratch of the bucket reduction (run sums, row / column totals, 256 slice sums) for three concurrent
        // sums: worst case over every window width msm_make_plan can pick
        s.sg_cap = 0;
        for (int c = 3; c <= 16; ++c) {
            MsmPlan p;
            p.nbits = 128; p.c = c; p.W = (128 + c - 1) / c;
            s.sg_cap = std::max(s.sg_cap, sg_work_entries(p));
        }
        s.sg_cap += 256;
        CK(dmalloc(s.sg_partial, 3 * s.sg_cap));
    }
    {
        const char* e = getenv("KZGB_SG_BATCH_MIN");
        if (e) s.sg_min = (size_t)strtoull(e, nullptr, 10);
    }
    CK(dmalloc(s.bucketsC, s.max_bucketsZ + 512));
    CK(dmalloc(s.winsums, 3 * KZ_MSM_MAX_WINDOWS));
    {   // chunk records for the balanced accumulation: worst case over the chunk-length schedule
        auto mn = [](size_t a, size_t b) { return a < b ? a : b; };
        size_t tmax = capZ / msm_chunk_len(capZ) + 1;
        size_t cands[3] = {mn(capZ, (size_t)1 << 21) / 16 + 1, mn(capZ, (size_t)1 << 20) / 8 + 1, mn(capZ, (size_t)1 << 18) / 4 + 1};
        for (size_t t : cands) if (t > tmax) tmax = t;
        tmax += 64;
        CK(dmalloc(s.recs.head, tmax)); CK(dmalloc(s.recs.tail, tmax));
        CK(dmalloc(s.recs.head_key, tmax)); CK(dmalloc(s.recs.tail_key, tmax)); CK(dmalloc(s.recs.head_flags, tmax));
        size_t tr = capR / msm_chunk_len(capR) + 1;
        size_t cr[3] = {mn(capR, (size_t)1 << 21) / 16 + 1, mn(capR, (size_t)1 << 20) / 8 + 1, mn(capR, (size_t)1 << 18) / 4 + 1};
        for (size_t t : cr) if (t > tr) tr = t;
        tr += 64;
        for (ChunkRecs* r : {&s.recsA, &s.recsB}) {
            CK(dmalloc(r->head, tr)); CK(dmalloc(r->tail, tr));
            CK(dmalloc(r->head_key, tr)); CK(dmalloc(r->tail_key, tr)); CK(dmalloc(r->head_flags, tr));
        }
    }
    CK(dmalloc(s.sums, 5)); CK(dmalloc(s.partial_dev, KZGB_PARTIAL_BYTES)); CK(dmalloc(s.partials_in, KZGB_PARTIAL_BYTES * 64));
    CK(dmalloc(s.scratch, 1 << 20));
    CK(dmalloc(s.result_dev, 4)); CK(dmalloc(s.lines, 2)); CK(dmalloc(s.g1_pt, 2)); CK(dmalloc(s.setup_status, 8));
    CK(cudaMallocHost((void**)&s.h_digests, 32 * nch)); CK(cudaMallocHost((void**)&s.h_small, 64 * sizeof(uint32_t)));
    CK(cudaMallocHost((void**)&s.h_partial, KZGB_PARTIAL_BYTES * 64));
    CK(dmalloc(s.mp_chain, KZ_MP_PAIRS)); CK(dmalloc(s.mp_tab, KZ_MP_PAIRS));
    CK(dmalloc(s.mp_terms, KZ_MP_PAIRS)); CK(dmalloc(s.mp_terms_in, 64 * KZ_MP_PAIRS)); CK(dmalloc(s.mp_coef, KZ_MP_PAIRS));
    CK(dmalloc(s.mp_part, mp_part_entries())); CK(dmalloc(s.mp_F, KZ_MP_ITERS));
    CK(cudaMallocHost((void**)&s.h_terms, (size_t)KZGB_TERMS_BYTES * 64));
    { const char* e = getenv("KZGB_CLASSIC"); s.classic = e && atoi(e) != 0; }
    // trusted setup: decompress + check on the device, precompute the G2 lines
    CK(cudaMemcpyAsync(s.scratch, g2m, 192, cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemcpyAsync(s.scratch + 256, g1m, 48, cudaMemcpyHostToDevice, s.stream));
    CK(cudaMemsetAsync(s.setup_status, 0, 8 * sizeof(int), s.stream));
    launch_g2_setup(s.stream, s.scratch, s.lines, s.setup_status);
    launch_g1_setup(s.stream, s.scratch + 256, s.g1_pt, s.setup_status);
    launch_mp_setup(s.stream, s.scratch, s.mp_chain, s.mp_tab, s.setup_status + 4);     // lines of [2^(4t)]G2, [2^(4t)][tau]G2
    int st[8];
    CK(cudaMemcpyAsync(st, s.setup_status, sizeof st, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    CK(cudaGetLastError());
    if (!(st[0] && st[1] && st[2] && st[4] && st[5] && st[6])) return KZGB_BADARGS;
    // powers of the 8192-th root of unity: cell batch and blob batch
    CK(dmalloc(s.cell_W, 8192));
    launch_cell_twiddles(s.stream, s.cell_W);
    if (n1 >= 64 && n2 >= 65) {
        // cell batch setup: 64 G1 monomials through K1 (decompress + subgroup), lines for (G2, [tau^64]G2), twiddles
        CK(dmalloc(s.cell_g1, 2 * 64)); CK(dmalloc(s.lines_cell, 2));
        CK(cudaMemcpyAsync(s.scratch, g1m, 48 * 64, cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(s.scratch + 4096, g2m, 96, cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemcpyAsync(s.scratch + 4096 + 96, g2m + 96 * 64, 96, cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemsetAsync(s.counters, 0, 8 * sizeof(uint32_t), s.stream));
        CK(cudaMemsetAsync(s.setup_status, 0, 8 * sizeof(int), s.stream));
        launch_decompress_points(s.stream, s.scratch, 64, s.cell_g1, s.k1_tmp, s.status, s.counters);
        launch_g2_setup(s.stream, s.scratch + 4096, s.lines_cell, s.setup_status);
        CK(dmalloc(s.mp_tab_cell, KZ_MP_PAIRS));
        launch_mp_setup(s.stream, s.scratch + 4096, s.mp_chain, s.mp_tab_cell, s.setup_status + 4);
        uint32_t cnt[2];
        uint8_t stat[64];
        CK(cudaMemcpyAsync(st, s.setup_status, sizeof st, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaMemcpyAsync(cnt, s.counters, sizeof cnt, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaMemcpyAsync(stat, s.status, 64, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaGetLastError());
        if (!(st[0] && st[1] && st[4] && st[5] && st[6]) || cnt[0]) return KZGB_BADARGS;
        s.cell_ready = true;
    }
    return KZGB_OK;
}

void slot_free(DeviceSlot& s) {
    cudaSetDevice(s.device);
    if (s.stream) cudaStreamSynchronize(s.stream);
    if (s.stream2) cudaStreamSynchronize(s.stream2);
    void* dev[] = {s.dC, s.dz, s.dy, s.dpi, s.pts, s.k1_tmp, s.status, s.counters, s.leaves, s.digests, s.root_words, s.r, s.rz,
                   s.partials, s.sum_ry, s.zs, s.sortR.keys, s.sortR.vals, s.sortR.keys_alt, s.sortR.vals_alt, s.sortR.bucket_start, s.sortR.count, s.sortR.cursor,
                   s.sortZ.count, s.sortZ.cursor, s.sortR.tile_sum, s.sortZ.tile_sum,
                   s.sortZ.keys, s.sortZ.vals, s.sortZ.keys_alt, s.sortZ.vals_alt, s.sortZ.bucket_start, s.bucketsA,
                   s.bucketsB, s.bucketsC, s.winsums, s.sums, s.partial_dev, s.partials_in,
                   s.scratch, s.result_dev, s.lines, s.g1_pt, s.setup_status, s.comb, s.recs.head, s.recs.tail,
                   s.recs.head_key, s.recs.tail_key, s.recs.head_flags, s.recsA.head, s.recsA.tail, s.recsA.head_key,
                   s.recsA.tail_key, s.recsA.head_flags, s.recsB.head, s.recsB.tail, s.recsB.head_key, s.recsB.tail_key,
                   s.recsB.head_flags, s.sg_partial, s.mp_chain, s.mp_tab, s.mp_tab_cell, s.mp_terms, s.mp_terms_in, s.mp_coef, s.mp_part, s.mp_F, s.cell_g1, s.lines_cell, s.cell_W, s.d_cells, s.d_ci, s.d_xi, s.cell_coefs, s.d_blobs, s.blob_leaves};
    for (void* p : dev) if (p) cudaFree(p);
    if (s.h_digests) cudaFreeHost(s.h_digests);
    if (s.h_small) cudaFreeHost(s.h_small);
    if (s.h_partial) cudaFreeHost(s.h_partial);
    if (s.h_terms) cudaFreeHost(s.h_terms);
    for (auto& g : s.tail_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    s.tail_graphs.clear();
    for (auto& e : s.ev) if (e) cudaEventDestroy(e);
    if (s.stream2) cudaStreamDestroy(s.stream2);
    if (s.stream3) cudaStreamDestroy(s.stream3);
    if (s.stream4) cudaStreamDestroy(s.stream4);
    if (s.stream5) cudaStreamDestroy(s.stream5);
    if (s.stream6) cudaStreamDestroy(s.stream6);
    if (s.stream) cudaStreamDestroy(s.stream);
}

MsmWorkspace make_ws(DeviceSlot& s, SortBuf& b, G1Xyzz* buckets) {
    MsmWorkspace ws;
    ws.keys = b.keys; ws.vals = b.vals; ws.keys_alt = b.keys_alt; ws.vals_alt = b.vals_alt;
    ws.capacity = b.capacity; ws.bucket_start = b.bucket_start; ws.count = b.count; ws.cursor = b.cursor; ws.tile_sum = b.tile_sum;
    ws.buckets = buckets;
    ws.winsums = s.winsums; ws.recs = s.recs;
    ws.sg_work = s.sg_partial; ws.slices = s.sg_partial + s.sg_cap - 256;
    ws.max_buckets = 0;
    return ws;
}
void save_ws(SortBuf& b, const MsmWorkspace& ws) {
    b.keys = ws.keys; b.vals = ws.vals; b.keys_alt = ws.keys_alt; b.vals_alt = ws.vals_alt;
}

// words (big-endian values) <-> bytes
void words_to_be(uint8_t* out, const uint32_t* w, size_t nw) {
    for (size_t i = 0; i < nw; ++i) { out[4 * i] = w[i] >> 24; out[4 * i + 1] = w[i] >> 16; out[4 * i + 2] = w[i] >> 8; out[4 * i + 3] = w[i]; }
}
void be_to_words(uint32_t* w, const uint8_t* in, size_t nw) {
    for (size_t i = 0; i < nw; ++i) w[i] = (uint32_t)in[4 * i] << 24 | in[4 * i + 1] << 16 | in[4 * i + 2] << 8 | in[4 * i + 3];
}

// Do the workspaces of this slot hold a shard of n proofs?  (They are sized at context creation from the plans of 64
// sampled sizes up to n_max plus slack; this is the exact check, done before anything is launched.)
bool shard_fits(const DeviceSlot& s, size_t n) {
    const MsmPlan pr = msm_make_plan(n, 128), pz = msm_make_plan(2 * (n + 1), 128);
    return (size_t)pr.W * n <= s.sortR.capacity && (size_t)pz.W * 2 * (n + 1) <= s.sortZ.capacity &&
           pr.total_buckets <= s.max_bucketsR + 512 && pz.total_buckets <= s.max_bucketsZ + 512 &&
           sg_work_entries(pr) + 256 <= s.sg_cap && sg_work_entries(pz) + 256 <= s.sg_cap;
}

// Phase 1: [H2D] -> leaf + chunk hashes -> digests D2H -> K1 decompress (left running).
// Returns after the digests are on the host.  `stream_override`: caller's stream or null.
kzgb_ret phase1(DeviceSlot& s, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi, size_t n,
                bool on_device, uint8_t* digests_out) {
    if (n == 0 || n > s.n_max) return KZGB_BADARGS;
    if (!shard_fits(s, n)) {                             // a workspace shortfall is the library's fault, not malformed input
        fprintf(stderr, "[kzgb200] workspace too small for a shard of %zu proofs (n_max %zu)\n", n, s.n_max);
        return KZGB_ERROR;
    }
    // 128-bit loads on the inputs: device-resident arrays must be 16-byte aligned (kzgb200.h)
    if (on_device && ((((uintptr_t)C) | ((uintptr_t)z) | ((uintptr_t)y) | ((uintptr_t)pi)) & 15u)) return KZGB_BADARGS;
    CK(cudaSetDevice(s.device));
    cudaStream_t st = s.stream, s2 = s.stream2;        // s2: high priority side stream
    CK(cudaEventRecord(s.ev[0], st));
    CK(cudaMemsetAsync(s.counters, 0, 8 * sizeof(uint32_t), st));
    CK(cudaEventRecord(s.ev[14], st));
    CK(cudaStreamWaitEvent(s2, s.ev[14], 0));
    s.sg_batch = s.sg_min && n >= s.sg_min && n >= 2;   // subgroup membership through the bucket slices of S1 and S3
    s.cur_n = n;
    s.have_sums = false;
    s.sums_pending = false;
    s.have_ab = false;
    s.head_mode = false;
    // Mid-size shards decompress the commitments first and record when they are done: sum r_i C_i (S1) then accumulates
    // and reduces UNDER K1 of the proofs on a high-priority stream, so its latency-bound tail and the idle partial waves
    // of its kernels cost nothing.  Measured on B200 (device-resident, ms without / with): n = 24576 2.54 / 2.55,
    // 49152 3.81 / 3.80, 65536 4.71 / 4.46, 81920 5.43 / 5.17, 98304 5.77 / 5.75, 131072 7.16 / 7.22, 2^18 12.57 / 12.88,
    // 2^20 44.65 / 44.81 -- it pays where K1 is a few waves long and the reduction tails are a large share; for larger
    // shards the mixed residency (232-register accumulation blocks displacing 142-register K1 blocks) costs more than the
    // tails save.  KZGB_S1_EARLY=0 switches it off, =3 forces it for every shard of >= 16384 proofs.
    s.s1_early = s.s1_early_enabled && s.sg_batch && ((n >= 57344 && n <= 90112) || (s.s1_early_force && n >= 16384));
    s.c_done_two = false;
    if (on_device) {
        s.cur_C = C; s.cur_z = z; s.cur_y = y; s.cur_pi = pi;
        CK(cudaEventRecord(s.ev[1], st));                // inputs already resident
        CK(cudaEventRecord(s.ev[13], s2));
        CK(cudaStreamWaitEvent(s2, s.ev[1], 0));
        // side stream: hashes start before K1 fills the SMs
        launch_leaf_hash(s2, C, z, y, pi, n, s.leaves, s.counters);
        launch_chunk_hash(s2, s.leaves, n, s.digests);
        if (s.sg_batch) {
            if (s.s1_early) {
                // commitments, then proofs on a second stream (its blocks fill the tail of the first launch)
                launch_decompress_sqrt_points(st, C, n, s.pts, s.status, s.counters);
                CK(cudaEventRecord(s.ev[26], st));
                CK(cudaStreamWaitEvent(s.stream5, s.ev[1], 0));
                launch_decompress_sqrt_points(s.stream5, pi, n, s.pts + 2 * n, s.status + n, s.counters);
                CK(cudaEventRecord(s.ev[25], s.stream5));
                CK(cudaStreamWaitEvent(st, s.ev[25], 0));
            } else {
                // ONE launch over both arrays: every launch boundary costs the idle tail of a partial wave
                launch_decompress_sqrt(st, C, pi, n, s.pts, s.status, s.counters);
            }
        } else {
            launch_decompress(st, C, pi, n, s.pts, s.k1_tmp, s.status, s.counters);
        }
    } else {
        // Host buffers (pinned or pageable -- the library stages nothing itself: a cudaMemcpyAsync from pageable memory
        // returns once the driver has staged the data, so every K1 launch is issued BEFORE the copies it does not depend
        // on).  All transfers go on ONE stream in the order K1 consumes them: C in growing pieces, pi, then z and y
        // (copies on the high-priority side stream overtake a copy queued on the main stream: measured).
        s.cur_C = s.dC; s.cur_z = s.dz; s.cur_y = s.dy; s.cur_pi = s.dpi;
        size_t cut[4] = {0, 0, 0, n};                    // C is copied (and decompressed) in the pieces [cut[j], cut[j+1])
        int ncut = 1;
        if (s.sg_batch && n >= 65536) {
            cudaPointerAttributes at;
            bool pageable = cudaPointerGetAttributes(&at, C) != cudaSuccess || at.type == cudaMemoryTypeUnregistered;
            cudaGetLastError();
            if (pageable) { cut[1] = (n / 16) & ~(size_t)127; cut[2] = (n / 4) & ~(size_t)127; ncut = 3; }
            else { cut[1] = (n / 8) & ~(size_t)127; cut[2] = n; ncut = 2; }
            s.head_mode = true;
        } else {
            cut[1] = n; cut[2] = n;
        }
        for (int j = 0; j < ncut; ++j) {
            const size_t a = cut[j], m = cut[j + 1] - a;
            CK(cudaMemcpyAsync(s.dC + 48 * a, C + 48 * a, 48 * m, cudaMemcpyHostToDevice, s2));
            cudaEvent_t e = j == 0 ? (s.head_mode ? s.ev[17] : s.ev[1]) : (j + 1 == ncut ? s.ev[1] : s.ev[23]);
            CK(cudaEventRecord(e, s2));
            if (s.head_mode) {
                // pieces alternate between two streams: the next piece starts in the idle tail of the previous one
                cudaStream_t ks = (j & 1) ? s.stream5 : st;
                if (j == 1) CK(cudaStreamWaitEvent(ks, s.ev[14], 0));          // after the counters' memset
                CK(cudaStreamWaitEvent(ks, e, 0));
                launch_decompress_sqrt_points(ks, s.dC + 48 * a, m, s.pts + 2 * a, s.status + a, s.counters);
            }
        }
        if (s.head_mode && s.s1_early) {                 // commitments decompressed: on the main stream, and on stream5 if a piece ran there
            CK(cudaEventRecord(s.ev[26], st));
            if (ncut >= 2) { CK(cudaEventRecord(s.ev[29], s.stream5)); s.c_done_two = true; }
        }
        CK(cudaMemcpyAsync(s.dpi, pi, 48 * n, cudaMemcpyHostToDevice, s2));
        CK(cudaEventRecord(s.ev[22], s2));               // pi resident
        CK(cudaStreamWaitEvent(st, s.ev[1], 0));
        CK(cudaStreamWaitEvent(st, s.ev[22], 0));
        if (s.head_mode) {
            cudaStream_t ks = (ncut & 1) ? s.stream5 : st;
            if (ks != st) CK(cudaStreamWaitEvent(ks, s.ev[22], 0));
            launch_decompress_sqrt_points(ks, s.dpi, n, s.pts + 2 * n, s.status + n, s.counters);
            CK(cudaEventRecord(s.ev[25], s.stream5));
            CK(cudaStreamWaitEvent(st, s.ev[25], 0));   // every piece has finished before the main stream goes on
        } else if (s.sg_batch && s.s1_early) {
            launch_decompress_sqrt_points(st, s.dC, n, s.pts, s.status, s.counters);
            CK(cudaEventRecord(s.ev[26], st));
            launch_decompress_sqrt_points(st, s.dpi, n, s.pts + 2 * n, s.status + n, s.counters);
        } else if (s.sg_batch) {
            launch_decompress_sqrt(st, s.dC, s.dpi, n, s.pts, s.status, s.counters);
        } else {
            launch_decompress(st, s.dC, s.dpi, n, s.pts, s.k1_tmp, s.status, s.counters);
        }
        CK(cudaMemcpyAsync(s.dz, z, 32 * n, cudaMemcpyHostToDevice, s2));
        CK(cudaMemcpyAsync(s.dy, y, 32 * n, cudaMemcpyHostToDevice, s2));
        CK(cudaEventRecord(s.ev[13], s2));               // all four arrays resident
        launch_leaf_hash(s2, s.dC, s.dz, s.dy, s.dpi, n, s.leaves, s.counters);
        launch_chunk_hash(s2, s.leaves, n, s.digests);
    }
    size_t nch = (n + KZGB_CHUNK - 1) / KZGB_CHUNK;
    CK(cudaMemcpyAsync(s.h_digests, s.digests, 32 * nch, cudaMemcpyDeviceToHost, s2));
    CK(cudaEventRecord(s.ev[2], s2));
    // the setup point G joins the GLV-split sum with scalar -(sum r_i y_i): point slot 2n; then phi(pi_i), phi(G)
    CK(cudaMemcpyAsync(s.pts + 2 * (2 * n), s.g1_pt, 2 * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    launch_endo_points(st, s.pts + 2 * n, n + 1, s.pts + 2 * (2 * n + 1));
    CK(cudaEventRecord(s.ev[3], st));
    CK(cudaEventSynchronize(s.ev[2]));
    words_to_be(digests_out, (const uint32_t*)s.h_digests, 8 * nch);
    return KZGB_OK;
}

// Phase 2: challenges, the three MSMs, partial.  Leaves the partial in s.h_partial[0..320) after sync.
// classic: Horner combine of the three sums and the 320-byte partial (shard-level ABI, stage exports); otherwise the
// shard's 66 pairing terms (mpair.cuh) are left in s.mp_terms and the sums are computed only if artefacts are requested.
kzgb_ret phase2(DeviceSlot& s, const uint8_t root[32], uint64_t global_offset, bool single, bool classic) {
    size_t n = s.cur_n;
    if (!n) return KZGB_BADARGS;
    CK(cudaSetDevice(s.device));
    cudaStream_t st = s.stream, s2 = s.stream2;
    // challenges, scalar products and both sorts depend only on the raw inputs: they run on the side
    // stream while K1 (decompress + subgroup checks) is still busy on the main stream
    be_to_words(s.h_small + 8, root, 8);
    CK(cudaMemcpyAsync(s.root_words, s.h_small + 8, 32, cudaMemcpyHostToDevice, s2));
    launch_challenges(s2, s.root_words, global_offset, s.cur_z, s.cur_y, n, single ? 1 : 0, s.r, s.rz, s.partials, s.sum_ry);
    launch_glv_split(s2, s.rz, n + 1, s.zs);       // r_i z_i and -(sum r_i y_i) -> 2(n+1) 128-bit scalars
    CK(cudaEventRecord(s.ev[4], s2));
    s.planR = msm_make_plan(n, 128);
    s.planZ = msm_make_plan(2 * (n + 1), 128);
    if (!shard_fits(s, n)) return KZGB_ERROR;            // checked in phase 1 already; never reached after a successful phase 1
    MsmWorkspace wr = make_ws(s, s.sortR, s.bucketsA), wz = make_ws(s, s.sortZ, s.bucketsC);
    msm_sort_stage(s2, s.planR, s.r, 4, n, wr);
    CK(cudaEventRecord(s.ev[28], s2));                   // digits of the r_i sorted: S1 can start once the commitments are decompressed
    msm_sort_stage(s2, s.planZ, s.zs, 4, 2 * (n + 1), wz);
    save_ws(s.sortR, wr); save_ws(s.sortZ, wz);
    CK(cudaEventRecord(s.ev[5], s2));
    CK(cudaStreamWaitEvent(st, s.ev[5], 0));
    // the three sums are independent from here to their per-window totals: run them on three streams
    MsmWorkspace wr2 = wr;
    wr.recs = s.recsA;
    wr2.buckets = s.bucketsB; wr2.recs = s.recsB;
    wr2.winsums = s.winsums + KZ_MSM_MAX_WINDOWS;
    wz.winsums = s.winsums + 2 * KZ_MSM_MAX_WINDOWS;
    wr2.sg_work = s.sg_partial + s.sg_cap; wr2.slices = wr2.sg_work + s.sg_cap - 256;
    wz.sg_work = s.sg_partial + 2 * s.sg_cap; wz.slices = wz.sg_work + s.sg_cap - 256;
    // Longest chain (S2', twice the points) first, all three sums at normal priority.
    const bool early = s.s1_early && !classic;
    cudaStream_t sS3 = s.stream3, sS1 = early ? s.stream6 : s.stream5;
    CK(cudaEventRecord(s.ev[11], st));
    CK(cudaStreamWaitEvent(sS3, s.ev[11], 0));
    CK(cudaStreamWaitEvent(s.stream4, s.ev[11], 0));
    if (early) {
        CK(cudaStreamWaitEvent(sS1, s.ev[28], 0));
        CK(cudaStreamWaitEvent(sS1, s.ev[26], 0));
        if (s.c_done_two) CK(cudaStreamWaitEvent(sS1, s.ev[29], 0));
    } else {
        CK(cudaStreamWaitEvent(sS1, s.ev[11], 0));
    }
    auto reduce = [&](cudaStream_t q, const MsmPlan& plan, MsmWorkspace& w, bool want_all) {
        if (classic) msm_window_sums_stage(q, plan, w, want_all); else msm_slices_stage(q, plan, w, want_all);
    };
    msm_accumulate_stage(s.stream4, s.planZ, s.pts + 2 * n, 2 * (n + 1), wz);    // S2' over pi_i, G and their phi images
    reduce(s.stream4, s.planZ, wz, false);
    CK(cudaEventRecord(s.ev[13], s.stream4));
    msm_accumulate_stage(sS3, s.planR, s.pts + 2 * n, n, wr2);                   // S3 over pi_i
    reduce(sS3, s.planR, wr2, s.sg_batch);
    CK(cudaEventRecord(s.ev[12], sS3));
    msm_accumulate_stage(sS1, s.planR, s.pts, n, wr);                            // S1 over C_i
    reduce(sS1, s.planR, wr, s.sg_batch);
    CK(cudaEventRecord(s.ev[16], sS1));
    if (s.sg_batch) {
        // batched subgroup check: 128 slice sums of the S3 buckets (all pi_i) and of the S1 buckets (all C_i)
        CK(cudaStreamWaitEvent(sS3, s.ev[16], 0));                               // slice sums of S1 are complete
        launch_sg_check(sS3, s.planR, wr2, wr, s.counters);
        // the verdict of the check travels on this stream: the main stream goes on to the pairing without it
        CK(cudaMemcpyAsync(s.h_small, s.counters, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, sS3));
        CK(cudaEventRecord(s.ev[9], sS3));
    }
    CK(cudaStreamWaitEvent(st, s.ev[12], 0));
    CK(cudaStreamWaitEvent(st, s.ev[13], 0));
    CK(cudaStreamWaitEvent(st, s.ev[16], 0));
    CK(cudaEventRecord(s.ev[6], st));
    s.ws_keep[0] = wr; s.ws_keep[1] = wr2; s.ws_keep[2] = wz;
    if (classic) {
        const MsmPlan* plans[3] = {&s.planR, &s.planR, &s.planZ};
        MsmWorkspace* wss[3] = {&wr, &wr2, &wz};
        G1Jac* outs[3] = {s.sums + 0, s.sums + 2, s.sums + 1};
        msm_combine_stage(st, plans, wss, outs, 3);
        launch_make_partial(st, s.sums + 0, s.sums + 1, s.sums + 2, s.sum_ry, s.partial_dev);
        CK(cudaMemcpyAsync(s.h_partial, s.partial_dev, KZGB_PARTIAL_BYTES, cudaMemcpyDeviceToHost, st));
        s.have_sums = true;
        s.sums_pending = false;
    } else {
        const MpSumDesc d1 = {wr.slices, wr.buckets, s.planR.c, s.planR.W, s.planR.nbits};
        const MpSumDesc d2 = {wz.slices, wz.buckets, s.planZ.c, s.planZ.W, s.planZ.nbits};
        const MpSumDesc d3 = {wr2.slices, wr2.buckets, s.planR.c, s.planR.W, s.planR.nbits};
        launch_mp_terms(st, d1, d2, d3, s.mp_terms);
        s.have_sums = false;
        s.sums_pending = true;
    }
    CK(cudaEventRecord(s.ev[7], st));
    if (!s.sg_batch) CK(cudaMemcpyAsync(s.h_small, s.counters, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    return KZGB_OK;
}
// S1, S2', S3 of the last batch by the Horner combine, from the slice sums that are still resident (artefacts only)
kzgb_ret deferred_sums(DeviceSlot& s) {
    if (!s.sums_pending) return KZGB_OK;
    CK(cudaSetDevice(s.device));
    cudaStream_t st = s.stream;
    const MsmPlan* plans[3] = {&s.planR, &s.planR, &s.planZ};
    MsmWorkspace* wss[3] = {&s.ws_keep[0], &s.ws_keep[1], &s.ws_keep[2]};
    G1Jac* outs[3] = {s.sums + 0, s.sums + 2, s.sums + 1};
    for (int j = 0; j < 3; ++j) msm_winsums_stage(st, *plans[j], *wss[j]);
    msm_combine_stage(st, plans, wss, outs, 3);
    s.sums_pending = false;
    s.have_sums = true;
    return KZGB_OK;
}

// After the main stream has been synchronised: if a slice sum of the batched subgroup check fell outside G1,
// find the offending points with the per-point chains (status bytes, counters[0]) -- the rare path.
kzgb_ret finish_subgroup(DeviceSlot& s) {
    if (!s.sg_batch) return KZGB_OK;
    CK(cudaEventSynchronize(s.ev[9]));                  // counters (incl. the check's verdict) are on the host
    if (!s.h_small[2]) return KZGB_OK;
    cudaStream_t st = s.stream;
    launch_subgroup_points(st, s.pts, 2 * s.cur_n, s.k1_tmp, s.status, s.counters);
    CK(cudaMemcpyAsync(s.h_small, s.counters, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    s.have_sums = false;
    s.sums_pending = false;
    return KZGB_OK;
}

kzgb_ret combine(DeviceSlot& s, const uint8_t* partials, int np, bool* ok) {
    if (np < 1 || np > 64) return KZGB_BADARGS;
    CK(cudaSetDevice(s.device));
    cudaStream_t st = s.stream;
    memcpy(s.h_partial, partials, (size_t)KZGB_PARTIAL_BYTES * np);
    CK(cudaMemcpyAsync(s.partials_in, s.h_partial, (size_t)KZGB_PARTIAL_BYTES * np, cudaMemcpyHostToDevice, st));
    launch_combine(st, s.partials_in, np, s.sums + 3, s.sum_ry);
    launch_pairing(st, s.lines, s.sums + 3, s.result_dev);
    CK(cudaMemcpyAsync(s.h_small + 16, s.result_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(s.ev[8], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    *ok = s.h_small[16] == 1;
    s.have_ab = true;
    return KZGB_OK;
}

// Pairing check on the terms of n_shards shards resident in terms_dev (device memory of s)
kzgb_ret mp_finish(DeviceSlot& s, const G1Xyzz* terms_dev, int n_shards, const G2Lines* tab, bool* ok) {
    cudaStream_t st = s.stream;
    cudaGraphExec_t exec = nullptr;
    if (s.use_graph) {
        for (auto& g : s.tail_graphs)
            if (g.terms == terms_dev && g.n_shards == n_shards && g.tab == tab) exec = g.exec;
        if (!exec) {
            // first use with these arguments: capture the sequence (nothing else is queued on a capturing stream meanwhile:
            // one in-flight call per slot), instantiate, keep
            cudaGraph_t graph = nullptr;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            launch_mp_coefs(st, terms_dev, n_shards, s.mp_coef);
            launch_mp_check(st, tab, s.mp_coef, s.mp_part, s.mp_F, s.result_dev);
            cudaMemcpyAsync(s.h_small + 16, s.result_dev, sizeof(int), cudaMemcpyDeviceToHost, st);
            CK(cudaStreamEndCapture(st, &graph));
            g_kzgb_launches.fetch_sub(4, std::memory_order_relaxed);   // the four kernels were captured, not launched
            CK(cudaGraphInstantiate(&exec, graph, 0));
            CK(cudaGraphDestroy(graph));
            s.tail_graphs.push_back({terms_dev, n_shards, tab, exec});
        }
        CK(cudaGraphLaunch(exec, st));
        g_kzgb_launches.fetch_add(4, std::memory_order_relaxed);   // k_mp_coefs, k_mp_lines, k_mp_merge, k_mp_check
    } else {
        launch_mp_coefs(st, terms_dev, n_shards, s.mp_coef);
        launch_mp_check(st, tab, s.mp_coef, s.mp_part, s.mp_F, s.result_dev);
        CK(cudaMemcpyAsync(s.h_small + 16, s.result_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaEventRecord(s.ev[8], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    *ok = s.h_small[16] == 1;
    return KZGB_OK;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return 0; }
    return ms;
}

// stages overlap (side stream): hash/challenges/sort are measured on the side stream, decompress and the
// rest on the main stream; "accumulate" starts when K1 has finished.  Valid after the slot's streams synced.
void fill_stage_ms(kzgb_artifacts& art, DeviceSlot& s0, float root_ms) {
    cudaEvent_t k1_start = s0.head_mode ? s0.ev[17] : s0.ev[1];
    art.stage_ms[0] = ev_ms(s0.ev[0], k1_start);
    art.stage_ms[2] = ev_ms(s0.ev[1], s0.ev[2]);
    art.stage_ms[1] = ev_ms(k1_start, s0.ev[3]);
    art.stage_ms[3] = root_ms;
    art.stage_ms[4] = ev_ms(s0.ev[2], s0.ev[4]);
    art.stage_ms[5] = ev_ms(s0.ev[4], s0.ev[5]);
    art.stage_ms[6] = ev_ms(s0.ev[3], s0.ev[6]);
    art.stage_ms[7] = ev_ms(s0.ev[6], s0.ev[7]);
    art.stage_ms[8] = 0;
    art.stage_ms[9] = ev_ms(s0.ev[0], s0.ev[7]);
}

kzgb_ret verify_common(bool* ok, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi, size_t n,
                       kzgb_ctx* ctx, bool on_device, bool single) {
    if (!ok) return KZGB_BADARGS;
    *ok = false;
    if (!ctx || !C || !z || !y || !pi || n == 0) return KZGB_BADARGS;
    size_t G = on_device ? 1 : ctx->slots.size();
    // contiguous shards, multiples of the hash chunk
    size_t nch = (n + KZGB_CHUNK - 1) / KZGB_CHUNK;
    if (G > nch) G = nch;
    std::vector<size_t> lo(G + 1);
    for (size_t g = 0; g <= G; ++g) { size_t c = nch * g / G * KZGB_CHUNK; lo[g] = c < n ? c : n; }
    lo[G] = n;
    const bool classic = ctx->slots[0].classic;
    ctx->terms_combined = false;
    std::vector<uint8_t> digests(32 * nch);
    std::vector<kzgb_ret> rcs(G, KZGB_OK);
    auto first_error = [&]() { for (kzgb_ret r : rcs) if (r) return r; return KZGB_OK; };
    // phase 1 on every shard at once (one host thread per device): copies, hashes, K1; returns with the digests
    ctx->pool.run(G, [&](size_t g) {
        size_t a = lo[g], m = lo[g + 1] - a;
        rcs[g] = phase1(ctx->slots[g], C + 48 * a, z + 32 * a, y + 32 * a, pi + 48 * a, m, on_device,
                        digests.data() + 32 * (a / KZGB_CHUNK));
    });
    if (kzgb_ret rc = first_error()) return rc;
    uint8_t root[32] = {0};
    auto t0 = std::chrono::steady_clock::now();
    if (!single) host_sha256_root(root, digests.data(), nch, n);
    float root_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    // One slot with the batched subgroup check: the pairing does not wait for the check's verdict (its |x|^2 chains
    // finish while the Miller loop runs); the counters are read after both.
    const bool lone = G == 1 && !classic;
    std::vector<uint32_t> badp_v(G, 0), bads_v(G, 0);
    ctx->pool.run(G, [&](size_t g) {
        DeviceSlot& s = ctx->slots[g];
        rcs[g] = phase2(s, root, lo[g], single, classic);
        if (rcs[g] || lone) return;
        auto body = [&]() -> kzgb_ret {
            if (!classic) CK(cudaMemcpyAsync(s.h_terms, s.mp_terms, sizeof(G1Xyzz) * KZ_MP_PAIRS, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaStreamSynchronize(s.stream));
            CK(cudaGetLastError());
            if (kzgb_ret frc = finish_subgroup(s)) return frc;
            badp_v[g] = s.h_small[0];
            bads_v[g] = s.h_small[1];
            return KZGB_OK;
        };
        rcs[g] = body();
    });
    if (kzgb_ret rc = first_error()) return rc;
    uint32_t badp = 0, bads = 0;
    for (size_t g = 0; g < G; ++g) { badp += badp_v[g]; bads += bads_v[g]; }
    kzgb_artifacts& art = ctx->art;
    memset(&art, 0, sizeof art);
    art.n = n;
    memcpy(art.root, root, 32);
    DeviceSlot& s0 = ctx->slots[0];
    CK(cudaSetDevice(s0.device));
    kzgb_ret rc = KZGB_OK;
    ctx->n_shards_last = (int)G;
    if (lone) {
        rc = mp_finish(s0, s0.mp_terms, 1, s0.mp_tab, ok);
        if (!rc) rc = finish_subgroup(s0);
        badp = s0.h_small[0];
        bads = s0.h_small[1];
    } else if (!(badp || bads)) {
        if (classic) {
            std::vector<uint8_t> parts(KZGB_PARTIAL_BYTES * G);
            for (size_t g = 0; g < G; ++g) memcpy(parts.data() + KZGB_PARTIAL_BYTES * g, ctx->slots[g].h_partial, KZGB_PARTIAL_BYTES);
            rc = combine(s0, parts.data(), (int)G, ok);
        } else {
            // "combined on the host": every shard's 66 terms came back through its pinned mailbox; the root device adds them
            for (size_t g = 0; g < G; ++g)
                CK(cudaMemcpyAsync(s0.mp_terms_in + g * KZ_MP_PAIRS, ctx->slots[g].h_terms, sizeof(G1Xyzz) * KZ_MP_PAIRS,
                                   cudaMemcpyHostToDevice, s0.stream));
            rc = mp_finish(s0, s0.mp_terms_in, (int)G, s0.mp_tab, ok);
            s0.have_ab = false;
            ctx->terms_combined = !rc;
            ctx->ab_terms = s0.mp_terms_in; ctx->ab_shards = (int)G; ctx->ab_gather_sum_ry = true;
        }
    }
    art.n_bad_points = badp;
    art.n_bad_scalars = bads;
    if (rc || badp || bads) {
        *ok = false;                                        // never leave a verdict next to an error code
        for (auto& s : ctx->slots) { s.have_sums = false; s.have_ab = false; s.sums_pending = false; }
        ctx->terms_combined = false;
        return rc ? rc : KZGB_BADARGS;
    }
    if (G > 1) { s0.have_sums = false; s0.sums_pending = false; }
    fill_stage_ms(art, s0, root_ms);
    art.stage_ms[8] = ev_ms(s0.ev[7], s0.ev[8]);
    art.stage_ms[9] = ev_ms(s0.ev[0], s0.ev[8]);
    return KZGB_OK;
}

// One batch on ONE workspace, start to verdict (the path a lane thread runs; same steps as verify_common with one slot).
kzgb_ret verify_on_slot(DeviceSlot& s, bool* ok, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi, size_t n,
                        bool on_device) {
    *ok = false;
    if (!C || !z || !y || !pi || n == 0) return KZGB_BADARGS;
    const size_t nch = (n + KZGB_CHUNK - 1) / KZGB_CHUNK;
    std::vector<uint8_t> digests(32 * nch);
    if (kzgb_ret rc = phase1(s, C, z, y, pi, n, on_device, digests.data())) return rc;
    uint8_t root[32];
    host_sha256_root(root, digests.data(), nch, n);
    if (kzgb_ret rc = phase2(s, root, 0, false, false)) return rc;
    kzgb_ret rc = mp_finish(s, s.mp_terms, 1, s.mp_tab, ok);
    if (!rc) rc = finish_subgroup(s);
    const bool bad = s.h_small[0] || s.h_small[1];
    s.have_sums = false; s.sums_pending = false; s.have_ab = false;
    if (rc || bad) { *ok = false; return rc ? rc : KZGB_BADARGS; }
    return KZGB_OK;
}

// EIP-4844 / c-kzg-4844 transcript: ONE challenge r = SHA-256 over the whole batch (host, serial by
// construction), coefficients r^i (255-bit).  All three sums are GLV-split 128-bit sums over 2n resp. 2(n+1) points:
//   S1  over  C | phi(C)                    scalars  k1(r^i) | k2(r^i)
//   S2' over  pi | G | phi(pi) | phi(G)     scalars  r^i z_i, -(sum r^i y_i)  split likewise
//   S3  over  the same 2(n+1) points        scalars  r^i, 0 (the setup point does not take part)
// Every point gets the deterministic per-point subgroup check: the coefficients r^i are functions of one value and
// not independent between points, so the argument of the batched check (128 independent coins per point) is not
// available here.  Needs 2(n+1) <= n_max.  `resident`: C, z, y, pi are already in s.dC .. s.dpi (blob path).
kzgb_ret eip_verify_on_slot(kzgb_ctx* ctx, DeviceSlot& s, bool* ok, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi,
                            size_t n, bool resident) {
    *ok = false;
    if (!C || !z || !y || !pi || n == 0 || 2 * (n + 1) > s.n_max || n >> KZ_EIP_POW_BITS) return KZGB_BADARGS;
    const MsmPlan pA = msm_make_plan(2 * n, 128), pZ = msm_make_plan(2 * (n + 1), 128);
    const size_t eZ = (size_t)pZ.W * 2 * (n + 1), eA = (size_t)pA.W * 2 * n;
    if (eA > s.sortR.capacity || eZ > s.sortR.capacity || eZ > s.sortZ.capacity || pA.total_buckets > s.max_bucketsR + 512 ||
        pZ.total_buckets > s.max_bucketsR + 512 || pZ.total_buckets > s.max_bucketsZ + 512 || sg_work_entries(pA) + 256 > s.sg_cap ||
        sg_work_entries(pZ) + 256 > s.sg_cap) {
        fprintf(stderr, "[kzgb200] workspace too small for an EIP-4844-mode batch of %zu proofs (n_max %zu)\n", n, s.n_max);
        return KZGB_ERROR;
    }
    CK(cudaSetDevice(s.device));
    cudaStream_t st = s.stream, sA = s.stream5, sZ = s.stream4;
    s.have_sums = false; s.sums_pending = false; s.have_ab = false; s.cur_n = 0; s.sg_batch = false;
    ctx->terms_combined = false;
    kzgb_artifacts& art = ctx->art;
    memset(&art, 0, sizeof art);
    art.n = n;
    CK(cudaEventRecord(s.ev[0], st));
    CK(cudaMemsetAsync(s.counters, 0, 8 * sizeof(uint32_t), st));
    if (!resident) {
        CK(cudaMemcpyAsync(s.dC, C, 48 * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s.dpi, pi, 48 * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s.dz, z, 32 * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s.dy, y, 32 * n, cudaMemcpyHostToDevice, st));
    }
    CK(cudaEventRecord(s.ev[1], st));
    // point layout (2 Fp each): C [0,n) | phi(C) [n,2n) | pi [2n,3n) | G 3n | phi(pi) [3n+1,4n+1) | phi(G) 4n+1
    Fp* pC = s.pts;
    Fp* pPi = s.pts + 2 * (2 * n);
    launch_decompress_points(st, s.dC, n, pC, s.k1_tmp, s.status, s.counters);
    launch_decompress_points(st, s.dpi, n, pPi, s.k1_tmp + 3 * n, s.status + n, s.counters);
    CK(cudaMemcpyAsync(pPi + 2 * n, s.g1_pt, 2 * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    launch_endo_points(st, pC, n, pC + 2 * n);
    launch_endo_points(st, pPi, n + 1, pPi + 2 * (n + 1));
    CK(cudaEventRecord(s.ev[3], st));
    // the transcript hash on the host while K1 runs, then powers and scalar products on the side stream
    uint8_t hash[32];
    host_eip4844_batch_hash(hash, C, z, y, pi, n);
    memcpy(art.root, hash, 32);
    memcpy(s.h_small + 8, hash, 32);
    cudaStream_t s2 = s.stream2;
    CK(cudaStreamWaitEvent(s2, s.ev[1], 0));
    uint8_t* hash_dev = s.scratch + 2048;
    Fr* table = (Fr*)(s.scratch + 4096);
    CK(cudaMemcpyAsync(hash_dev, s.h_small + 8, 32, cudaMemcpyHostToDevice, s2));
    uint32_t* rpow = s.rz;                          // 8 (n+1) limbs
    uint32_t* rz = s.rz + 8 * (n + 1);              // 8 (n+1) limbs: needs 16 (n+1) <= 8 (n_max+1)
    launch_eip_scalars(s2, hash_dev, table, s.dz, s.dy, n, s.root_words, rpow, rz, s.partials, s.sum_ry, s.counters);
    uint32_t* zsA = s.r;                            // 4 * 2n limbs <= 4 n_max
    uint32_t* zs3 = s.zs;                           // 4 * 2(n+1)
    uint32_t* zsZ = s.zs + 4 * 2 * (n + 1);         // 4 * 2(n+1): needs 16 (n+1) <= 8 (n_max+1)
    launch_glv_split(s2, rpow, n, zsA);
    launch_glv_split(s2, rpow, n + 1, zs3);
    launch_glv_split(s2, rz, n + 1, zsZ);
    CK(cudaEventRecord(s.ev[4], s2));
    CK(cudaStreamWaitEvent(st, s.ev[4], 0));
    CK(cudaEventRecord(s.ev[11], st));              // points and scalars ready
    CK(cudaStreamWaitEvent(sA, s.ev[11], 0));
    CK(cudaStreamWaitEvent(sZ, s.ev[11], 0));
    MsmWorkspace w1 = make_ws(s, s.sortR, s.bucketsA), w3 = make_ws(s, s.sortR, s.bucketsB), w2 = make_ws(s, s.sortZ, s.bucketsC);
    w1.recs = s.recsA; w3.recs = s.recsB;
    w3.sg_work = s.sg_partial + s.sg_cap; w3.slices = w3.sg_work + s.sg_cap - 256;
    w2.sg_work = s.sg_partial + 2 * s.sg_cap; w2.slices = w2.sg_work + s.sg_cap - 256;
    // S2' on its own stream; S1 and then S3 share the second sort buffer, one after the other on one stream
    msm_sort_stage(sZ, pZ, zsZ, 4, 2 * (n + 1), w2);
    save_ws(s.sortZ, w2);
    msm_accumulate_stage(sZ, pZ, pPi, 2 * (n + 1), w2);
    msm_slices_stage(sZ, pZ, w2, false);
    CK(cudaEventRecord(s.ev[13], sZ));
    msm_sort_stage(sA, pA, zsA, 4, 2 * n, w1);
    msm_accumulate_stage(sA, pA, pC, 2 * n, w1);
    msm_slices_stage(sA, pA, w1, false);
    w3.keys = w1.keys; w3.vals = w1.vals; w3.keys_alt = w1.keys_alt; w3.vals_alt = w1.vals_alt;
    msm_sort_stage(sA, pZ, zs3, 4, 2 * (n + 1), w3);
    save_ws(s.sortR, w3);
    msm_accumulate_stage(sA, pZ, pPi, 2 * (n + 1), w3);
    msm_slices_stage(sA, pZ, w3, false);
    CK(cudaEventRecord(s.ev[12], sA));
    CK(cudaStreamWaitEvent(st, s.ev[12], 0));
    CK(cudaStreamWaitEvent(st, s.ev[13], 0));
    CK(cudaEventRecord(s.ev[6], st));
    const MpSumDesc d1 = {w1.slices, w1.buckets, pA.c, pA.W, pA.nbits};
    const MpSumDesc d2 = {w2.slices, w2.buckets, pZ.c, pZ.W, pZ.nbits};
    const MpSumDesc d3 = {w3.slices, w3.buckets, pZ.c, pZ.W, pZ.nbits};
    launch_mp_terms(st, d1, d2, d3, s.mp_terms);
    CK(cudaEventRecord(s.ev[7], st));
    CK(cudaMemcpyAsync(s.h_small, s.counters, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    kzgb_ret rc = mp_finish(s, s.mp_terms, 1, s.mp_tab, ok);
    art.n_bad_points = s.h_small[0];
    art.n_bad_scalars = s.h_small[1];
    if (rc || s.h_small[0] || s.h_small[1]) { *ok = false; return rc ? rc : KZGB_BADARGS; }
    ctx->terms_combined = true;
    ctx->ab_terms = s.mp_terms; ctx->ab_shards = 1; ctx->ab_gather_sum_ry = false;
    fill_stage_ms(art, s, 0.0f);
    art.stage_ms[8] = ev_ms(s.ev[7], s.ev[8]);
    art.stage_ms[9] = ev_ms(s.ev[0], s.ev[8]);
    return KZGB_OK;
}

void lane_main(Lane* L) {
    std::unique_lock<std::mutex> lk(L->mu);
    for (;;) {
        L->cv.wait(lk, [L] { return L->state == Lane::RUNNING || L->quit; });
        if (L->quit) return;
        lk.unlock();
        bool ok = false;
        kzgb_ret rc = verify_on_slot(L->slot, &ok, L->C, L->z, L->y, L->pi, L->n, L->on_device);
        lk.lock();
        L->rc = rc; L->ok = ok;
        L->state = Lane::DONE;
        L->cv.notify_all();
    }
}

}  // namespace

extern "C" {

const char* kzgb_version(void) { return "kzgb200-cuda-sm100a 0.1"; }

kzgb_ret kzgb_ctx_create(kzgb_ctx** out, const uint8_t* g1m, size_t n1, const uint8_t* g2m, size_t n2, const int* devices,
                         int n_devices, size_t n_max) {
    KZ_API_GUARD;
    if (!out || !g1m || !g2m || n1 < 1 || n2 < 2 || n_max < 1 || n_devices < 0 || n_devices > 64) return KZGB_BADARGS;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        fprintf(stderr, "[kzgb200] no CUDA device: this library has no CPU fallback\n");
        return KZGB_ERROR;
    }
    kzgb_ctx* c = new (std::nothrow) kzgb_ctx();
    if (!c) return KZGB_MALLOC;
    int nd = n_devices > 0 && devices ? n_devices : 1;
    c->slots.resize(nd);
    for (int i = 0; i < nd; ++i) {
        int dev = (n_devices > 0 && devices) ? devices[i] : 0;
        if (dev < 0 || dev >= ndev) { kzgb_ctx_free(c); return KZGB_BADARGS; }
        kzgb_ret rc = slot_init(c->slots[i], dev, n_max, g1m, n1, g2m, n2);
        if (rc) { kzgb_ctx_free(c); return rc; }
    }
    memset(&c->art, 0, sizeof c->art);
    c->setup_g1.assign(g1m, g1m + 48 * n1);
    c->setup_g2.assign(g2m, g2m + 96 * n2);
    if (nd > 1) c->pool.start((size_t)nd - 1);
    c->launches_at_create = g_kzgb_launches.load();
    *out = c;
    return KZGB_OK;
}
void kzgb_ctx_free(kzgb_ctx* c) {
    KZ_API_GUARD;
    if (!c) return;
    c->pool.stop();
    for (auto& L : c->lanes) {
        { std::lock_guard<std::mutex> lk(L->mu); L->quit = true; }
        L->cv.notify_all();
        if (L->th.joinable()) L->th.join();
        slot_free(L->slot);
    }
    c->lanes.clear();
    for (auto& s : c->slots) slot_free(s);
    delete c;
}

kzgb_ret verify_kzg_proof(bool* ok, const uint8_t C[48], const uint8_t z[32], const uint8_t y[32], const uint8_t pi[48],
                          kzgb_ctx* ctx) {
    KZ_API_GUARD;
    return verify_common(ok, C, z, y, pi, 1, ctx, false, true);
}
kzgb_ret verify_kzg_proof_batch(bool* ok, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi, size_t n,
                                kzgb_ctx* ctx) {
    KZ_API_GUARD;
    return verify_common(ok, C, z, y, pi, n, ctx, false, false);
}
kzgb_ret verify_kzg_proof_batch_device(bool* ok, const uint8_t* dC, const uint8_t* dz, const uint8_t* dy, const uint8_t* dpi,
                                       size_t n, kzgb_ctx* ctx, void* stream) {
    KZ_API_GUARD;
    // inputs may have been produced on the caller's stream: order our stream after it
    if (ctx && !ctx->slots.empty()) {
        DeviceSlot& s = ctx->slots[0];
        CK(cudaSetDevice(s.device));
        CK(cudaEventRecord(s.ev[15], (cudaStream_t)stream));
        CK(cudaStreamWaitEvent(s.stream, s.ev[15], 0));
    }
    return verify_common(ok, dC, dz, dy, dpi, n, ctx, true, false);
}

// ---- submit / wait: up to `depth` batches in flight on device 0 of the context, one workspace + host thread each
kzgb_ret kzgb_pipeline_init(kzgb_ctx* ctx, int depth) {
    KZ_API_GUARD;
    if (!ctx || depth < 1 || depth > 8) return KZGB_BADARGS;
    std::lock_guard<std::mutex> g(ctx->lane_mu);
    if ((int)ctx->lanes.size() == depth) return KZGB_OK;
    for (auto& L : ctx->lanes) {
        std::lock_guard<std::mutex> lk(L->mu);
        if (L->state != Lane::IDLE) return KZGB_BADARGS;           // batches in flight: wait for them first
    }
    while ((int)ctx->lanes.size() > depth) {
        auto& L = ctx->lanes.back();
        { std::lock_guard<std::mutex> lk(L->mu); L->quit = true; }
        L->cv.notify_all();
        L->th.join();
        slot_free(L->slot);
        ctx->lanes.pop_back();
    }
    while ((int)ctx->lanes.size() < depth) {
        std::unique_ptr<Lane> L(new (std::nothrow) Lane());
        if (!L) return KZGB_MALLOC;
        const DeviceSlot& s0 = ctx->slots[0];
        kzgb_ret rc = slot_init(L->slot, s0.device, s0.n_max, ctx->setup_g1.data(), 1, ctx->setup_g2.data(), 2);
        if (rc) { slot_free(L->slot); return rc; }
        L->slot.sg_min = s0.sg_min;
        Lane* raw = L.get();
        L->th = std::thread(lane_main, raw);
        ctx->lanes.push_back(std::move(L));
    }
    return KZGB_OK;
}
kzgb_ret verify_kzg_proof_batch_submit(uint64_t* ticket_out, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi,
                                       size_t n, int inputs_on_device, kzgb_ctx* ctx) {
    if (!ticket_out || !ctx || !C || !z || !y || !pi || n == 0) return KZGB_BADARGS;
    std::lock_guard<std::mutex> g(ctx->lane_mu);
    if (ctx->lanes.empty()) return KZGB_BADARGS;                   // kzgb_pipeline_init first
    const uint64_t t = ctx->next_ticket;
    Lane* L = ctx->lanes[t % ctx->lanes.size()].get();
    if (n > L->slot.n_max) return KZGB_BADARGS;
    {
        std::lock_guard<std::mutex> lk(L->mu);
        if (L->state != Lane::IDLE) return KZGB_BADARGS;           // `depth` batches already in flight: wait for ticket t - depth
        L->C = C; L->z = z; L->y = y; L->pi = pi; L->n = n; L->on_device = inputs_on_device != 0;
        L->ticket = t;
        L->state = Lane::RUNNING;
    }
    L->cv.notify_all();
    ctx->next_ticket = t + 1;
    *ticket_out = t;
    return KZGB_OK;
}
kzgb_ret verify_kzg_proof_batch_wait(bool* ok, uint64_t ticket, kzgb_ctx* ctx) {
    if (!ok) return KZGB_BADARGS;
    *ok = false;
    if (!ctx) return KZGB_BADARGS;
    Lane* L;
    {
        std::lock_guard<std::mutex> g(ctx->lane_mu);
        if (ctx->lanes.empty() || ticket >= ctx->next_ticket) return KZGB_BADARGS;
        L = ctx->lanes[ticket % ctx->lanes.size()].get();
    }
    std::unique_lock<std::mutex> lk(L->mu);
    if (L->state == Lane::IDLE || L->ticket != ticket) return KZGB_BADARGS;       // unknown or already collected
    L->cv.wait(lk, [L] { return L->state == Lane::DONE; });
    const kzgb_ret rc = L->rc;
    *ok = rc == KZGB_OK && L->ok;
    L->state = Lane::IDLE;
    return rc;
}

kzgb_ret kzgb_shard_phase1(kzgb_ctx* ctx, int slot, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi,
                           size_t n_local, int on_device, void* stream, uint8_t* digests_out, uint32_t* n_bad_out) {
    KZ_API_GUARD;
    if (!ctx || slot < 0 || slot >= (int)ctx->slots.size() || !C || !z || !y || !pi || !digests_out) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[slot];
    if (on_device) {
        CK(cudaSetDevice(s.device));
        CK(cudaEventRecord(s.ev[15], (cudaStream_t)stream));
        CK(cudaStreamWaitEvent(s.stream, s.ev[15], 0));
    }
    kzgb_ret rc = phase1(s, C, z, y, pi, n_local, on_device != 0, digests_out);
    if (n_bad_out) *n_bad_out = 0;      // malformed elements are reported by phase 2 (K1 is still running)
    return rc;
}
kzgb_ret kzgb_fs_root(uint8_t root_out[32], const uint8_t* chunk_digests, size_t n_chunks, uint64_t n_total) {
    if (!root_out || !chunk_digests) return KZGB_BADARGS;
    host_sha256_root(root_out, chunk_digests, n_chunks, n_total);
    return KZGB_OK;
}
kzgb_ret kzgb_shard_phase2(kzgb_ctx* ctx, int slot, const uint8_t root[32], uint64_t global_offset, void*,
                           uint8_t partial_out[KZGB_PARTIAL_BYTES]) {
    KZ_API_GUARD;
    if (!ctx || slot < 0 || slot >= (int)ctx->slots.size() || !root || !partial_out) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[slot];
    kzgb_ret rc = phase2(s, root, global_offset, false, true);
    if (rc) return rc;
    CK(cudaStreamSynchronize(s.stream));
    CK(cudaGetLastError());
    rc = finish_subgroup(s);
    if (rc) return rc;
    memcpy(partial_out, s.h_partial, KZGB_PARTIAL_BYTES);
    fill_stage_ms(ctx->art, s, 0.0f);
    ctx->art.n_bad_points = s.h_small[0];
    ctx->art.n_bad_scalars = s.h_small[1];
    return (s.h_small[0] || s.h_small[1]) ? KZGB_BADARGS : KZGB_OK;
}
// Shard-level entry points of the Horner-free pairing check: the shard's 66 pairing terms (mpair.cuh) instead of the
// 320-byte partial -- no serial Horner chain on any shard.  The call returns as soon as the terms are on the host; the
// verdict of the shard's input validation (malformed points / scalars, batched subgroup check) is collected by
// kzgb_shard_finish, which can be called after the terms have been handed on.
kzgb_ret kzgb_shard_phase2_terms(kzgb_ctx* ctx, int slot, const uint8_t root[32], uint64_t global_offset, void*,
                                 uint8_t terms_out[KZGB_TERMS_BYTES]) {
    KZ_API_GUARD;
    if (!ctx || slot < 0 || slot >= (int)ctx->slots.size() || !root || !terms_out) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[slot];
    kzgb_ret rc = phase2(s, root, global_offset, false, false);
    if (rc) return rc;
    uint8_t* wire = s.scratch + 16384;
    launch_mp_terms_to_wire(s.stream, s.mp_terms, s.sum_ry, wire);
    CK(cudaMemcpyAsync(s.h_terms, wire, KZGB_TERMS_BYTES, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaEventRecord(s.ev[7], s.stream));
    CK(cudaStreamSynchronize(s.stream));
    CK(cudaGetLastError());
    memcpy(terms_out, s.h_terms, KZGB_TERMS_BYTES);
    fill_stage_ms(ctx->art, s, 0.0f);
    return KZGB_OK;
}
kzgb_ret kzgb_shard_finish(kzgb_ctx* ctx, int slot, uint32_t* n_bad_points, uint32_t* n_bad_scalars) {
    KZ_API_GUARD;
    if (!ctx || slot < 0 || slot >= (int)ctx->slots.size()) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[slot];
    if (!s.cur_n) return KZGB_BADARGS;
    CK(cudaSetDevice(s.device));
    CK(cudaStreamSynchronize(s.stream));                // per-point path: the counters were copied on the main stream
    if (kzgb_ret rc = finish_subgroup(s)) return rc;
    ctx->art.n_bad_points = s.h_small[0];
    ctx->art.n_bad_scalars = s.h_small[1];
    if (n_bad_points) *n_bad_points = s.h_small[0];
    if (n_bad_scalars) *n_bad_scalars = s.h_small[1];
    return (s.h_small[0] || s.h_small[1]) ? KZGB_BADARGS : KZGB_OK;
}
kzgb_ret kzgb_combine_verify_terms(kzgb_ctx* ctx, const uint8_t* terms, int n_shards, bool* ok) {
    KZ_API_GUARD;
    if (!ctx || !terms || !ok) return KZGB_BADARGS;
    *ok = false;
    if (n_shards < 1 || n_shards > 64) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[0];
    CK(cudaSetDevice(s.device));
    cudaStream_t st = s.stream;
    // second half of the pinned mailbox: the first record may still hold this slot's own outgoing terms
    uint8_t* stage = s.h_terms;
    memcpy(stage, terms, (size_t)KZGB_TERMS_BYTES * n_shards);
    uint8_t* wire = s.scratch + 65536;
    uint32_t* bad = (uint32_t*)(s.scratch + 32768);
    CK(cudaMemsetAsync(bad, 0, sizeof(uint32_t), st));
    CK(cudaMemcpyAsync(wire, stage, (size_t)KZGB_TERMS_BYTES * n_shards, cudaMemcpyHostToDevice, st));
    launch_mp_terms_from_wire(st, wire, n_shards, s.mp_terms_in, s.sum_ry, bad);
    CK(cudaMemcpyAsync(s.h_small + 20, bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    kzgb_ret rc = mp_finish(s, s.mp_terms_in, n_shards, s.mp_tab, ok);
    s.have_sums = false; s.sums_pending = false; s.have_ab = false;
    ctx->terms_combined = false;
    if (rc || s.h_small[20]) { *ok = false; return rc ? rc : KZGB_BADARGS; }
    ctx->terms_combined = true;
    ctx->ab_terms = s.mp_terms_in; ctx->ab_shards = n_shards; ctx->ab_gather_sum_ry = false;
    ctx->art.stage_ms[8] = ev_ms(s.ev[7], s.ev[8]);
    return KZGB_OK;
}
kzgb_ret kzgb_combine_verify(kzgb_ctx* ctx, const uint8_t* partials, int n_partials, bool* ok) {
    KZ_API_GUARD;
    if (!ctx || !partials || !ok) return KZGB_BADARGS;
    *ok = false;
    DeviceSlot& s = ctx->slots[0];
    if (n_partials != 1) s.have_sums = false;
    return combine(s, partials, n_partials, ok);
}

// ---- blob batch: z and y of every blob on the device, left in s.dz / s.dy
#define KZ_BLOB_STAGE 256
static kzgb_ret blob_zy(DeviceSlot& s, const uint8_t* blobs, const uint8_t* comms, const uint8_t* z_in, size_t m, uint32_t* n_bad) {
    if (m == 0 || m > s.n_max) return KZGB_BADARGS;
    CK(cudaSetDevice(s.device));
    cudaStream_t st = s.stream;
    if (!s.d_blobs) {
        CK(dmalloc(s.d_blobs, (size_t)2 * KZ_BLOB_STAGE * KZGB_BLOB_BYTES));        // two staging buffers
        CK(dmalloc(s.blob_leaves, (size_t)2 * KZ_BLOB_STAGE * 128 * 8));
    }
    uint32_t* bad_dev = (uint32_t*)(s.scratch + 4096);
    CK(cudaMemsetAsync(bad_dev, 0, sizeof(uint32_t), st));
    if (comms) CK(cudaMemcpyAsync(s.dC, comms, 48 * m, cudaMemcpyHostToDevice, st));
    if (z_in) CK(cudaMemcpyAsync(s.dz, z_in, 32 * m, cudaMemcpyHostToDevice, st));
    // double buffering: the copy stream fills one staging buffer while the kernels work on the other
    cudaStream_t cp = s.stream2;
    CK(cudaEventRecord(s.ev[14], st));
    CK(cudaStreamWaitEvent(cp, s.ev[14], 0));
    size_t chunk = 0;
    for (size_t done = 0; done < m; done += KZ_BLOB_STAGE, ++chunk) {
        size_t k = m - done < KZ_BLOB_STAGE ? m - done : KZ_BLOB_STAGE;
        int b = (int)(chunk & 1);
        uint8_t* buf = s.d_blobs + (size_t)b * KZ_BLOB_STAGE * KZGB_BLOB_BYTES;
        uint32_t* leaves = s.blob_leaves + (size_t)b * KZ_BLOB_STAGE * 128 * 8;
        if (chunk >= 2) CK(cudaStreamWaitEvent(cp, s.ev[20 + b], 0));              // buffer b is free again
        CK(cudaMemcpyAsync(buf, blobs + (size_t)KZGB_BLOB_BYTES * done, (size_t)KZGB_BLOB_BYTES * k, cudaMemcpyHostToDevice, cp));
        CK(cudaEventRecord(s.ev[18 + b], cp));
        CK(cudaStreamWaitEvent(st, s.ev[18 + b], 0));
        if (!z_in) launch_blob_challenges(st, buf, s.dC + 48 * done, k, leaves, s.dz + 32 * done);
        launch_blob_eval(st, s.cell_W, buf, s.dz + 32 * done, k, s.dy + 32 * done, bad_dev);
        CK(cudaEventRecord(s.ev[20 + b], st));
    }
    CK(cudaMemcpyAsync(s.h_small + 24, bad_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    *n_bad = s.h_small[24];
    return KZGB_OK;
}
kzgb_ret kzgb_blob_challenges_evals(uint8_t* z_out, uint8_t* y_out, const uint8_t* blobs, const uint8_t* comms, size_t m,
                                    kzgb_ctx* ctx) {
    KZ_API_GUARD;
    if (!z_out || !y_out || !blobs || !comms || !ctx || m == 0) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[0];
    uint32_t bad = 0;
    kzgb_ret rc = blob_zy(s, blobs, comms, nullptr, m, &bad);
    if (rc) return rc;
    CK(cudaMemcpyAsync(z_out, s.dz, 32 * m, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(y_out, s.dy, 32 * m, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    memset(&ctx->art, 0, sizeof ctx->art);
    ctx->art.n = m;
    ctx->art.n_bad_scalars = bad;
    return bad ? KZGB_BADARGS : KZGB_OK;
}
kzgb_ret kzgb_blob_eval(uint8_t* y_out, const uint8_t* blobs, const uint8_t* z_in, size_t m, kzgb_ctx* ctx) {
    KZ_API_GUARD;
    if (!y_out || !blobs || !z_in || !ctx || m == 0) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[0];
    uint32_t bad = 0;
    kzgb_ret rc = blob_zy(s, blobs, nullptr, z_in, m, &bad);
    if (rc) return rc;
    CK(cudaMemcpyAsync(y_out, s.dy, 32 * m, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    return bad ? KZGB_BADARGS : KZGB_OK;
}
kzgb_ret verify_blob_kzg_proof_batch(bool* ok, const uint8_t* blobs, const uint8_t* comms, const uint8_t* proofs, size_t m,
                                     kzgb_ctx* ctx) {
    KZ_API_GUARD;
    if (!ok) return KZGB_BADARGS;
    *ok = false;
    if (!ctx || !blobs || !comms || !proofs || m == 0) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[0];
    uint32_t bad = 0;
    kzgb_ret rc = blob_zy(s, blobs, comms, nullptr, m, &bad);
    if (rc) return rc;
    if (bad) {
        memset(&ctx->art, 0, sizeof ctx->art);
        ctx->art.n = m;
        ctx->art.n_bad_scalars = bad;
        return KZGB_BADARGS;
    }
    CK(cudaMemcpyAsync(s.dpi, proofs, 48 * m, cudaMemcpyHostToDevice, s.stream));
    // the plain batch on device-resident (C, z, y, pi); one proof is still a batch of one (challenge r_0 from the root)
    return verify_common(ok, s.dC, s.dz, s.dy, s.dpi, m, ctx, true, false);
}

// ---- EIP-4844 / c-kzg-4844 transcript mode
kzgb_ret verify_kzg_proof_batch_eip4844(bool* ok, const uint8_t* C, const uint8_t* z, const uint8_t* y, const uint8_t* pi, size_t n,
                                        kzgb_ctx* ctx) {
    KZ_API_GUARD;
    if (!ok) return KZGB_BADARGS;
    *ok = false;
    if (!ctx) return KZGB_BADARGS;
    return eip_verify_on_slot(ctx, ctx->slots[0], ok, C, z, y, pi, n, false);
}
kzgb_ret verify_blob_kzg_proof_batch_eip4844(bool* ok, const uint8_t* blobs, const uint8_t* comms, const uint8_t* proofs, size_t m,
                                             kzgb_ctx* ctx) {
    KZ_API_GUARD;
    if (!ok) return KZGB_BADARGS;
    *ok = false;
    if (!ctx || !blobs || !comms || !proofs || m == 0) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[0];
    if (2 * (m + 1) > s.n_max) return KZGB_BADARGS;
    CK(cudaSetDevice(s.device));
    // compute_challenge of every blob: a flat SHA-256 over 128 KiB + 80 bytes -- host threads, SHA extensions
    std::vector<uint8_t> zs(32 * m), ys(32 * m);
    {
        unsigned nt = std::thread::hardware_concurrency();
        nt = nt < 1 ? 1 : (nt > 16 ? 16 : nt);
        if (nt > m) nt = (unsigned)m;
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t)
            th.emplace_back([&, t] {
                for (size_t j = t; j < m; j += nt) host_eip4844_blob_hash(zs.data() + 32 * j, blobs + (size_t)KZGB_BLOB_BYTES * j, comms + 48 * j);
            });
        for (auto& t : th) t.join();
    }
    CK(cudaMemcpyAsync(s.dz, zs.data(), 32 * m, cudaMemcpyHostToDevice, s.stream));
    launch_eip_reduce_be(s.stream, s.dz, m);                       // z = hash mod r
    CK(cudaMemcpyAsync(zs.data(), s.dz, 32 * m, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    uint32_t bad = 0;
    kzgb_ret rc = blob_zy(s, blobs, comms, zs.data(), m, &bad);   // y = p(z) on the device (barycentric); C, z, y stay resident
    if (rc) return rc;
    if (bad) {
        memset(&ctx->art, 0, sizeof ctx->art);
        ctx->art.n = m;
        ctx->art.n_bad_scalars = bad;
        return KZGB_BADARGS;
    }
    CK(cudaMemcpyAsync(ys.data(), s.dy, 32 * m, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaMemcpyAsync(s.dpi, proofs, 48 * m, cudaMemcpyHostToDevice, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    return eip_verify_on_slot(ctx, s, ok, comms, zs.data(), ys.data(), proofs, m, true);
}
// stage export: z_out, y_out of the EIP-4844 blob path (compute_challenge, evaluate_polynomial_in_evaluation_form)
kzgb_ret kzgb_blob_challenges_evals_eip4844(uint8_t* z_out, uint8_t* y_out, const uint8_t* blobs, const uint8_t* comms, size_t m,
                                            kzgb_ctx* ctx) {
    KZ_API_GUARD;
    if (!z_out || !y_out || !blobs || !comms || !ctx || m == 0) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[0];
    if (m > s.n_max) return KZGB_BADARGS;
    CK(cudaSetDevice(s.device));
    for (size_t j = 0; j < m; ++j) host_eip4844_blob_hash(z_out + 32 * j, blobs + (size_t)KZGB_BLOB_BYTES * j, comms + 48 * j);
    CK(cudaMemcpyAsync(s.dz, z_out, 32 * m, cudaMemcpyHostToDevice, s.stream));
    launch_eip_reduce_be(s.stream, s.dz, m);
    CK(cudaMemcpyAsync(z_out, s.dz, 32 * m, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    uint32_t bad = 0;
    kzgb_ret rc = blob_zy(s, blobs, nullptr, z_out, m, &bad);
    if (rc) return rc;
    CK(cudaMemcpyAsync(y_out, s.dy, 32 * m, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    return bad ? KZGB_BADARGS : KZGB_OK;
}

// c-kzg-4844 trusted_setup.txt: "<n_g1>\n<n_g2>\n", n_g1 lines of G1 points in Lagrange form (96 hex digits; used only by
// provers -- skipped), n_g2 lines of G2 monomials (192 hex digits) and, in files written for EIP-7594, n_g1 more lines
// of G1 monomials.  The verifier needs [tau^0]G1 (the generator: taken from the monomial section when present), and
// [tau^j]G2; with the monomial section it also gets the 64 G1 monomials of the cell batch.
static int hex_nibble(int c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1; }
static bool read_hex_token(FILE* f, std::vector<uint8_t>& out, size_t nbytes) {
    int c;
    do { c = fgetc(f); } while (c == ' ' || c == '\n' || c == '\r' || c == '\t');
    if (c == EOF) return false;
    out.clear();
    for (size_t i = 0; i < nbytes; ++i) {
        int hi = hex_nibble(c), lo = hex_nibble(fgetc(f));
        if (hi < 0 || lo < 0) return false;
        out.push_back((uint8_t)(hi << 4 | lo));
        c = fgetc(f);
    }
    if (c != EOF && c != '\n' && c != '\r' && c != ' ' && c != '\t') return false;       // token longer than expected
    return true;
}
kzgb_ret kzgb_load_trusted_setup_file(kzgb_ctx** out, const char* path, const int* devices, int n_devices, size_t n_max) {
    if (!out || !path) return KZGB_BADARGS;
    FILE* f = fopen(path, "r");
    if (!f) return KZGB_BADARGS;
    unsigned long n1 = 0, n2 = 0;
    std::vector<uint8_t> g1, g2, tok;
    bool okf = fscanf(f, "%lu %lu", &n1, &n2) == 2 && n1 >= 1 && n1 <= (1ul << 20) && n2 >= 2 && n2 <= 4096;
    for (unsigned long i = 0; okf && i < n1; ++i) okf = read_hex_token(f, tok, 48);                 // Lagrange section: not used
    for (unsigned long i = 0; okf && i < n2; ++i) { okf = read_hex_token(f, tok, 96); if (okf) g2.insert(g2.end(), tok.begin(), tok.end()); }
    if (okf) {
        unsigned long got = 0;
        while (got < n1 && read_hex_token(f, tok, 48)) { g1.insert(g1.end(), tok.begin(), tok.end()); ++got; }
        if (got != 0 && got != n1) okf = false;                    // a truncated monomial section
        if (got == 0) {
            // no monomial section: [tau^0]G1 is the group generator
            static const char* gen = "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb";
            for (int i = 0; i < 48; ++i) g1.push_back((uint8_t)(hex_nibble(gen[2 * i]) << 4 | hex_nibble(gen[2 * i + 1])));
        }
    }
    fclose(f);
    if (!okf) return KZGB_BADARGS;
    return kzgb_ctx_create(out, g1.data(), g1.size() / 48, g2.data(), g2.size() / 96, devices, n_devices, n_max);
}

kzgb_ret kzgb_g1_decompress_batch(uint8_t* affine_out, uint8_t* status_out, const uint8_t* in, size_t m, kzgb_ctx* ctx) {
    KZ_API_GUARD;
    if (!affine_out || !status_out || !in || !ctx) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[0];
    CK(cudaSetDevice(s.device));
    uint8_t* d_be = nullptr;
    size_t slice = s.n_max;
    CK(cudaMalloc((void**)&d_be, 96 * (slice + 1)));
    for (size_t done = 0; done < m; done += slice) {
        // K1 takes two input arrays of h points each: split the slice in two halves of one buffer
        size_t k = m - done < slice ? m - done : slice;
        size_t h = (k + 1) / 2;
        CK(cudaMemsetAsync(s.dC + 48 * k, 0, 48, s.stream));             // pad slot when k is odd
        CK(cudaMemcpyAsync(s.dC, in + 48 * done, 48 * k, cudaMemcpyHostToDevice, s.stream));
        CK(cudaMemsetAsync(s.counters, 0, 8 * sizeof(uint32_t), s.stream));
        launch_decompress(s.stream, s.dC, s.dC + 48 * h, h, s.pts, s.k1_tmp, s.status, s.counters);
        launch_points_to_be(s.stream, s.pts, k, d_be);
        CK(cudaMemcpyAsync(affine_out + 96 * done, d_be, 96 * k, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaMemcpyAsync(status_out + done, s.status, k, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
    }
    CK(cudaFree(d_be));
    CK(cudaGetLastError());
    s.have_sums = false;
    s.sums_pending = false;
    return KZGB_OK;
}

kzgb_ret kzgb_fs_challenges(uint8_t root_out[32], uint8_t* r_out, const uint8_t* C, const uint8_t* z, const uint8_t* y,
                            const uint8_t* pi, size_t n, kzgb_ctx* ctx) {
    KZ_API_GUARD;
    if (!root_out || !r_out || !C || !z || !y || !pi || !ctx || n == 0) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[0];
    if (n > s.n_max) return KZGB_BADARGS;
    size_t nch = (n + KZGB_CHUNK - 1) / KZGB_CHUNK;
    std::vector<uint8_t> dig(32 * nch);
    kzgb_ret rc = phase1(s, C, z, y, pi, n, false, dig.data());
    if (rc) return rc;
    host_sha256_root(root_out, dig.data(), nch, n);
    be_to_words(s.h_small + 8, root_out, 8);
    CK(cudaMemcpyAsync(s.root_words, s.h_small + 8, 32, cudaMemcpyHostToDevice, s.stream));
    uint8_t* d_r = nullptr;
    CK(cudaMalloc((void**)&d_r, 16 * n));
    launch_r_only(s.stream, s.root_words, n, d_r);
    CK(cudaMemcpyAsync(r_out, d_r, 16 * n, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    CK(cudaFree(d_r));
    CK(cudaGetLastError());
    return KZGB_OK;
}

kzgb_ret kzgb_g1_msm(uint8_t affine_out[96], const uint8_t* points_affine, const uint8_t* scalars, size_t m, int nbits,
                     kzgb_ctx* ctx) {
    KZ_API_GUARD;
    if (!affine_out || !points_affine || !scalars || !ctx || (nbits != 255 && nbits != 128)) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[0];
    if (m > s.n_max) return KZGB_BADARGS;
    if (m == 0) { memset(affine_out, 0, 96); return KZGB_OK; }
    CK(cudaSetDevice(s.device));
    cudaStream_t st = s.stream;
    uint8_t *d_pts = nullptr, *d_sc = nullptr;
    CK(cudaMalloc((void**)&d_pts, 96 * m)); CK(cudaMalloc((void**)&d_sc, 32 * m));
    CK(cudaMemcpyAsync(d_pts, points_affine, 96 * m, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_sc, scalars, 32 * m, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(s.counters, 0, 8 * sizeof(uint32_t), st));
    launch_points_from_be(st, d_pts, m, s.pts, s.counters);
    launch_scalars_from_be(st, d_sc, m, s.rz, s.counters);
    // 255-bit scalars are GLV-split: 2m points (P, phi(P)) with 128-bit scalars (k1, k2)
    const bool glv = nbits == 255;
    size_t mm = glv ? 2 * m : m;
    MsmPlan plan = msm_make_plan(mm, 128);
    if ((size_t)plan.W * mm > s.sortZ.capacity || plan.total_buckets > s.max_bucketsZ + 512 || sg_work_entries(plan) + 256 > s.sg_cap) {
        cudaFree(d_pts); cudaFree(d_sc);
        return KZGB_ERROR;
    }
    MsmWorkspace ws = make_ws(s, s.sortZ, s.bucketsC);
    CK(cudaEventRecord(s.ev[0], st));
    if (glv) {
        launch_glv_split(st, s.rz, m, s.zs);
        launch_endo_points(st, s.pts, m, s.pts + 2 * m);
        msm_sort_stage(st, plan, s.zs, 4, mm, ws);
    } else {
        msm_sort_stage(st, plan, s.rz, 8, mm, ws);
    }
    save_ws(s.sortZ, ws);
    CK(cudaEventRecord(s.ev[1], st));
    msm_accumulate_stage(st, plan, s.pts, mm, ws);
    CK(cudaEventRecord(s.ev[2], st));
    msm_reduce_stage(st, plan, ws, s.sums + 0);
    CK(cudaEventRecord(s.ev[3], st));
    launch_jac_to_affine_be(st, s.sums + 0, 1, s.scratch);
    CK(cudaMemcpyAsync(s.h_partial, s.scratch, 96, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(s.h_small, s.counters, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(d_pts)); CK(cudaFree(d_sc));
    CK(cudaGetLastError());
    s.have_sums = false;
    s.sums_pending = false;
    ctx->msm_ms[0] = ev_ms(s.ev[0], s.ev[1]);
    ctx->msm_ms[1] = ev_ms(s.ev[1], s.ev[2]);
    ctx->msm_ms[2] = ev_ms(s.ev[2], s.ev[3]);
    ctx->msm_ms[3] = ev_ms(s.ev[0], s.ev[3]);
    if (s.h_small[0] || s.h_small[1]) return KZGB_BADARGS;
    if (nbits == 128) {
        for (size_t i = 0; i < m; ++i)
            for (int k = 0; k < 16; ++k)
                if (scalars[32 * i + k]) return KZGB_BADARGS;
    }
    memcpy(affine_out, s.h_partial, 96);
    return KZGB_OK;
}
kzgb_ret kzgb_g1_msm_times(float ms_out[4], kzgb_ctx* ctx) {
    if (!ms_out || !ctx) return KZGB_BADARGS;
    memcpy(ms_out, ctx->msm_ms, sizeof ctx->msm_ms);
    return KZGB_OK;
}

// Cell batch (BASELINE.json config[4]: "on 8xB200").  The openings shard like the plain batch: contiguous ranges, multiples
// of the 128-leaf hash chunk, one per device of the context.  Everything is linear in the openings, so a shard needs no
// exchange beyond the chunk digests and the root: it runs K1 on its proofs and on ALL nc commitments, its openings' iNTTs,
// its PARTIAL column sums S_j and commitment weights w_i, and one A-side sum over [its proofs | all commitments | [tau^j]G1]
// with those partial scalars plus the B-side sum over its proofs -- the shards' 66 pairing terms add up to the batch's.
struct CellShard {
    size_t lo = 0, m = 0, nc = 0;
    bool sg = false;
    MsmPlan planA, planB;
};
// phase 1 of one shard: copies, leaf + chunk hashes (digests on the host when it returns), K1 left running
static kzgb_ret cell_phase1(DeviceSlot& s, CellShard& sh, const uint8_t* comms, const uint32_t* ci, const uint32_t* xi,
                            const uint8_t* cells, const uint8_t* proofs, uint8_t* digests_out) {
    const size_t m = sh.m, nc = sh.nc, lo = sh.lo;
    CK(cudaSetDevice(s.device));
    cudaStream_t st = s.stream;
    if (m > s.cell_cap) {
        for (void* p : {(void*)s.d_cells, (void*)s.d_ci, (void*)s.d_xi, (void*)s.cell_coefs}) if (p) cudaFree(p);
        s.d_cells = nullptr; s.d_ci = s.d_xi = nullptr; s.cell_coefs = nullptr; s.cell_cap = 0;
        CK(dmalloc(s.d_cells, 2048 * m)); CK(dmalloc(s.d_ci, m)); CK(dmalloc(s.d_xi, m)); CK(dmalloc(s.cell_coefs, 64 * m));
        s.cell_cap = m;
    }
    s.have_sums = false; s.have_ab = false; s.sums_pending = false; s.cur_n = 0;
    CK(cudaEventRecord(s.ev[0], st));
    CK(cudaMemsetAsync(s.counters, 0, 8 * sizeof(uint32_t), st));
    CK(cudaMemcpyAsync(s.dC, comms, 48 * nc, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s.dpi, proofs + 48 * lo, 48 * m, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s.d_ci, ci + lo, 4 * m, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s.d_xi, xi + lo, 4 * m, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s.d_cells, cells + 2048 * lo, 2048 * m, cudaMemcpyHostToDevice, st));
    // Fiat-Shamir on the side stream: cell leaves -> chunk digests -> host root (binds the commitments as well)
    CK(cudaEventRecord(s.ev[1], st));
    CK(cudaStreamWaitEvent(s.stream2, s.ev[1], 0));
    launch_cell_leaf_hash(s.stream2, s.d_ci, s.d_xi, s.d_cells, s.dpi, m, s.leaves);
    launch_chunk_hash(s.stream2, s.leaves, m, s.digests);
    const size_t nch = (m + KZGB_CHUNK - 1) / KZGB_CHUNK;
    CK(cudaMemcpyAsync(s.h_digests, s.digests, 32 * nch, cudaMemcpyDeviceToHost, s.stream2));
    CK(cudaEventRecord(s.ev[2], s.stream2));
    // K1 on proofs and commitments; the 64 setup monomials follow them in the point array.  Large shards prove subgroup
    // membership of the proofs on the bucket slices of the B-side sum (sum r_k pi_k), like the plain batch; the few
    // commitments get the per-point check on their own stream (three serial chains of pure latency)
    sh.sg = s.sg_min && m >= s.sg_min && m >= 2;
    s.sg_batch = sh.sg;
    CK(cudaStreamWaitEvent(s.stream4, s.ev[1], 0));
    launch_decompress_points(s.stream4, s.dC, nc, s.pts + 2 * m, s.k1_tmp + 3 * m, s.status + m, s.counters);
    CK(cudaEventRecord(s.ev[13], s.stream4));
    if (sh.sg) launch_decompress_sqrt_points(st, s.dpi, m, s.pts, s.status, s.counters);
    else launch_decompress_points(st, s.dpi, m, s.pts, s.k1_tmp, s.status, s.counters);
    CK(cudaStreamWaitEvent(st, s.ev[13], 0));
    CK(cudaMemcpyAsync(s.pts + 2 * (m + nc), s.cell_g1, 2 * 64 * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    CK(cudaEventSynchronize(s.ev[2]));
    words_to_be(digests_out, (const uint32_t*)s.h_digests, 8 * nch);
    return KZGB_OK;
}
// phase 2 of one shard: scalars, both sums, the shard's 66 pairing terms in s.mp_terms (classic: A, B in s.sums[3..4])
static kzgb_ret cell_phase2(DeviceSlot& s, const CellShard& sh, const uint8_t root[32]) {
    const size_t m = sh.m, nc = sh.nc, M = m + nc + 64;
    CK(cudaSetDevice(s.device));
    cudaStream_t st = s.stream;
    be_to_words(s.h_small + 8, root, 8);
    CK(cudaMemcpyAsync(s.root_words, s.h_small + 8, 32, cudaMemcpyHostToDevice, st));
    // per opening: r_k, r_k h^64, r_k * interpolation coefficients; then column sums and commitment weights
    launch_cell_scalars(st, s.cell_W, s.root_words, s.d_ci, s.d_xi, (uint32_t)nc, s.d_cells, m, (uint64_t)sh.lo, s.cell_coefs, s.r, s.rz,
                        s.counters);
    launch_cell_reductions(st, s.cell_coefs, s.d_ci, s.r, m, (uint32_t)nc, s.rz + 8 * m, s.rz + 8 * (m + nc));
    CK(cudaMemsetAsync(s.sum_ry, 0, 8 * sizeof(uint32_t), st));
    CK(cudaEventRecord(s.ev[11], st));
    // B-side: sum r_k pi_k, 128-bit scalars, on a side stream beside the A-side sum
    const MsmPlan& planB = sh.planB;
    const MsmPlan& planA = sh.planA;
    MsmWorkspace wsB = make_ws(s, s.sortR, s.bucketsA);
    {
        cudaStream_t sb = s.stream3;
        CK(cudaStreamWaitEvent(sb, s.ev[11], 0));
        wsB.recs = s.recsA;
        wsB.winsums = s.winsums + KZ_MSM_MAX_WINDOWS;
        wsB.sg_work = s.sg_partial + s.sg_cap; wsB.slices = wsB.sg_work + s.sg_cap - 256;
        msm_sort_stage(sb, planB, s.r, 4, m, wsB);
        save_ws(s.sortR, wsB);
        msm_accumulate_stage(sb, planB, s.pts, m, wsB);
        if (s.classic) msm_window_sums_stage(sb, planB, wsB, sh.sg); else msm_slices_stage(sb, planB, wsB, sh.sg);
        CK(cudaEventRecord(s.ev[12], sb));
        if (sh.sg) {
            launch_sg_check(sb, planB, wsB, wsB, s.counters, 1);
            CK(cudaMemcpyAsync(s.h_small, s.counters, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, sb));
            CK(cudaEventRecord(s.ev[9], sb));
        }
    }
    // A-side: sum (r_k h^64) pi_k + sum w_i C_i - sum S_j [tau^j]G1, 255-bit scalars, GLV-split
    size_t mm = 2 * M;
    MsmWorkspace wsA = make_ws(s, s.sortZ, s.bucketsC);
    launch_glv_split(st, s.rz, M, s.zs);
    launch_endo_points(st, s.pts, M, s.pts + 2 * M);
    msm_sort_stage(st, planA, s.zs, 4, mm, wsA);
    save_ws(s.sortZ, wsA);
    msm_accumulate_stage(st, planA, s.pts, mm, wsA);
    if (s.classic) msm_window_sums_stage(st, planA, wsA, false); else msm_slices_stage(st, planA, wsA, false);
    CK(cudaStreamWaitEvent(st, s.ev[12], 0));
    if (s.classic) {
        const MsmPlan* plans[2] = {&planA, &planB};
        MsmWorkspace* wss[2] = {&wsA, &wsB};
        G1Jac* outs[2] = {s.sums + 0, s.sums + 2};
        msm_combine_stage(st, plans, wss, outs, 2);
        launch_set_ab(st, s.sums + 0, s.sums + 2, s.sums + 3);
    } else {
        // A-side sum and -(B-side sum) as 2 x 33 pairing terms against the multiples of (G2, [tau^64]G2)
        const MpSumDesc dA = {wsA.slices, wsA.buckets, planA.c, planA.W, planA.nbits};
        const MpSumDesc dNone = {nullptr, nullptr, 0, 0, -1};
        const MpSumDesc dB = {wsB.slices, wsB.buckets, planB.c, planB.W, planB.nbits};
        launch_mp_terms(st, dA, dNone, dB, s.mp_terms);
    }
    CK(cudaEventRecord(s.ev[7], st));
    if (!sh.sg) CK(cudaMemcpyAsync(s.h_small, s.counters, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    return KZGB_OK;
}
// input-validation verdict of one shard (after its main stream has been synchronised); on a failed slice check the
// per-point chains name the proofs
static kzgb_ret cell_finish(DeviceSlot& s, const CellShard& sh) {
    if (!sh.sg) return KZGB_OK;
    CK(cudaSetDevice(s.device));
    CK(cudaEventSynchronize(s.ev[9]));
    if (s.h_small[2]) {
        launch_subgroup_points(s.stream, s.pts, sh.m, s.k1_tmp, s.status, s.counters);
        CK(cudaMemcpyAsync(s.h_small, s.counters, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaGetLastError());
    }
    s.sg_batch = false;
    return KZGB_OK;
}
kzgb_ret verify_cell_kzg_proof_batch(bool* ok, const uint8_t* comms, size_t nc, const uint32_t* ci, const uint32_t* xi,
                                     const uint8_t* cells, const uint8_t* proofs, size_t m, kzgb_ctx* ctx) {
    KZ_API_GUARD;
    if (!ok) return KZGB_BADARGS;
    *ok = false;
    if (!ctx || !comms || !ci || !xi || !cells || !proofs || m == 0 || nc == 0 || nc > 0xFFFFFFFFull) return KZGB_BADARGS;
    // shards: contiguous, multiples of the hash chunk; a shard below 1024 openings is all fixed latency, so small batches
    // use fewer devices
    const size_t nch = (m + KZGB_CHUNK - 1) / KZGB_CHUNK;
    size_t G = ctx->slots.size();
    const bool classic = ctx->slots[0].classic;
    if (classic) G = 1;
    while (G > 1 && (m / G < 1024 || G > nch)) --G;
    std::vector<CellShard> sh(G);
    for (size_t g = 0; g < G; ++g) {
        size_t a = nch * g / G * KZGB_CHUNK, b = g + 1 == G ? m : nch * (g + 1) / G * KZGB_CHUNK;
        sh[g].lo = a; sh[g].m = b - a; sh[g].nc = nc;
        DeviceSlot& s = ctx->slots[g];
        const size_t M = sh[g].m + nc + 64;                   // points of the A-side sum: proofs | commitments | [tau^j]G1
        if (!s.cell_ready || M > s.n_max) return KZGB_BADARGS;
        // both sums' plans against the workspaces, before anything is launched
        sh[g].planB = msm_make_plan(sh[g].m, 128);
        sh[g].planA = msm_make_plan(2 * M, 128);
        const MsmPlan &pB = sh[g].planB, &pA = sh[g].planA;
        if ((size_t)pB.W * sh[g].m > s.sortR.capacity || pB.total_buckets > s.max_bucketsR + 512 || sg_work_entries(pB) + 256 > s.sg_cap ||
            (size_t)pA.W * 2 * M > s.sortZ.capacity || pA.total_buckets > s.max_bucketsZ + 512 || sg_work_entries(pA) + 256 > s.sg_cap) {
            fprintf(stderr, "[kzgb200] workspace too small for a cell batch shard of %zu openings (n_max %zu)\n", sh[g].m, s.n_max);
            return KZGB_ERROR;
        }
    }
    kzgb_artifacts& art = ctx->art;
    memset(&art, 0, sizeof art);
    art.n = m;
    ctx->terms_combined = false;
    ctx->n_shards_last = (int)G;
    std::vector<uint8_t> dig(32 * nch);
    std::vector<kzgb_ret> rcs(G, KZGB_OK);
    auto first_error = [&]() { for (kzgb_ret r : rcs) if (r) return r; return KZGB_OK; };
    ctx->pool.run(G, [&](size_t g) {
        rcs[g] = cell_phase1(ctx->slots[g], sh[g], comms, ci, xi, cells, proofs, dig.data() + 32 * (sh[g].lo / KZGB_CHUNK));
    });
    if (kzgb_ret rc = first_error()) return rc;
    uint8_t root[32];
    host_sha256_cell_root(root, comms, nc, dig.data(), nch, m);
    memcpy(art.root, root, 32);
    DeviceSlot& s0 = ctx->slots[0];
    std::vector<uint32_t> badp(G, 0), bads(G, 0);
    ctx->pool.run(G, [&](size_t g) {
        DeviceSlot& s = ctx->slots[g];
        rcs[g] = cell_phase2(s, sh[g], root);
        if (rcs[g] || G == 1) return;
        auto body = [&]() -> kzgb_ret {
            CK(cudaMemcpyAsync(s.h_terms, s.mp_terms, sizeof(G1Xyzz) * KZ_MP_PAIRS, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaStreamSynchronize(s.stream));
            CK(cudaGetLastError());
            if (kzgb_ret frc = cell_finish(s, sh[g])) return frc;
            badp[g] = s.h_small[0]; bads[g] = s.h_small[1];
            return KZGB_OK;
        };
        rcs[g] = body();
    });
    if (kzgb_ret rc = first_error()) return rc;
    CK(cudaSetDevice(s0.device));
    cudaStream_t st = s0.stream;
    kzgb_ret rc = KZGB_OK;
    if (G == 1) {
        if (classic) launch_pairing(st, s0.lines_cell, s0.sums + 3, s0.result_dev);
        else {
            launch_mp_coefs(st, s0.mp_terms, 1, s0.mp_coef);
            launch_mp_check(st, s0.mp_tab_cell, s0.mp_coef, s0.mp_part, s0.mp_F, s0.result_dev);
        }
        CK(cudaMemcpyAsync(s0.h_small + 16, s0.result_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(s0.ev[8], st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        rc = cell_finish(s0, sh[0]);                      // the batched check ran beside the pairing
        badp[0] = s0.h_small[0]; bads[0] = s0.h_small[1];
        *ok = s0.h_small[16] == 1;
    } else {
        uint32_t anyb = 0;
        for (size_t g = 0; g < G; ++g) anyb |= badp[g] | bads[g];
        if (!anyb) {
            for (size_t g = 0; g < G; ++g)
                CK(cudaMemcpyAsync(s0.mp_terms_in + g * KZ_MP_PAIRS, ctx->slots[g].h_terms, sizeof(G1Xyzz) * KZ_MP_PAIRS,
                                   cudaMemcpyHostToDevice, st));
            rc = mp_finish(s0, s0.mp_terms_in, (int)G, s0.mp_tab_cell, ok);
        }
    }
    uint32_t tp = 0, ts = 0;
    for (size_t g = 0; g < G; ++g) { tp += badp[g]; ts += bads[g]; }
    art.n_bad_points = tp;
    art.n_bad_scalars = ts;
    art.stage_ms[9] = ev_ms(s0.ev[0], s0.ev[8]);
    if (rc || tp || ts) { *ok = false; return rc ? rc : KZGB_BADARGS; }
    if (classic) s0.have_ab = true;
    else {
        ctx->terms_combined = true;
        ctx->ab_terms = G == 1 ? s0.mp_terms : s0.mp_terms_in;
        ctx->ab_shards = (int)G;
        ctx->ab_gather_sum_ry = false;
    }
    return KZGB_OK;
}

kzgb_ret kzgb_pairing_check(bool* ok, const uint8_t A_affine[96], const uint8_t B_affine[96], kzgb_ctx* ctx) {
    KZ_API_GUARD;
    if (!ok || !A_affine || !B_affine || !ctx) return KZGB_BADARGS;
    *ok = false;
    DeviceSlot& s = ctx->slots[0];
    CK(cudaSetDevice(s.device));
    cudaStream_t st = s.stream;
    memcpy(s.h_partial, A_affine, 96);
    memcpy(s.h_partial + 96, B_affine, 96);
    CK(cudaMemcpyAsync(s.scratch, s.h_partial, 192, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(s.counters, 0, 8 * sizeof(uint32_t), st));
    launch_points_from_be(st, s.scratch, 2, s.pts, s.counters);       // validates range + curve equation
    CK(cudaMemcpyAsync(s.h_small, s.counters, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    launch_points_jac_from_be(st, s.scratch, 2, s.sums + 3);
    launch_pairing(st, s.lines, s.sums + 3, s.result_dev);
    CK(cudaMemcpyAsync(s.h_small + 16, s.result_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    s.have_sums = false;
    s.sums_pending = false;
    if (s.h_small[0]) return KZGB_BADARGS;
    *ok = s.h_small[16] == 1;
    return KZGB_OK;
}

kzgb_ret kzgb_last_artifacts(kzgb_ctx* ctx, kzgb_artifacts* out) {
    KZ_API_GUARD;
    if (!ctx || !out) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[0];
    if (s.sums_pending) { if (kzgb_ret rc = deferred_sums(s)) return rc; }
    if (ctx->terms_combined) {
        // several shards, or a cell batch: A and B by Horner over the combined pairing terms (cold path)
        if (ctx->ab_gather_sum_ry) {
            for (int g = 0; g < ctx->ab_shards; ++g) {
                DeviceSlot& sg = ctx->slots[g];
                CK(cudaSetDevice(sg.device));
                CK(cudaMemcpy(s.h_small + 32, sg.sum_ry, 32, cudaMemcpyDeviceToHost));
                CK(cudaSetDevice(s.device));
                CK(cudaMemcpy(s.scratch + 8192 + 32 * g, s.h_small + 32, 32, cudaMemcpyHostToDevice));
            }
            CK(cudaSetDevice(s.device));
            launch_fr_sum(s.stream, (const uint32_t*)(s.scratch + 8192), ctx->ab_shards, s.sum_ry);
        }
        CK(cudaSetDevice(s.device));
        launch_mp_ab(s.stream, ctx->ab_terms, ctx->ab_shards, s.sums + 3);
        s.have_ab = true;
        s.have_sums = false;
        ctx->terms_combined = false;
    }
    if (s.have_sums) {
        CK(cudaSetDevice(s.device));
        launch_artifacts(s.stream, s.sums + 0, s.sums + 1, s.sums + 2, s.sum_ry, s.g1_pt, s.scratch);
        CK(cudaMemcpyAsync(s.h_partial, s.scratch, 512, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaGetLastError());
        memcpy(ctx->art.S1, s.h_partial, 96); memcpy(ctx->art.S2, s.h_partial + 96, 96);
        memcpy(ctx->art.S3, s.h_partial + 192, 96); memcpy(ctx->art.A, s.h_partial + 288, 96);
        memcpy(ctx->art.B, s.h_partial + 384, 96); memcpy(ctx->art.sum_ry, s.h_partial + 480, 32);
    } else if (s.have_ab) {
        // several shards or a cell batch: only the pairing inputs (and the total sum r_i y_i) exist on this device
        CK(cudaSetDevice(s.device));
        launch_jac_to_affine_be(s.stream, s.sums + 3, 2, s.scratch);
        launch_fr_to_be(s.stream, s.sum_ry, s.scratch + 192);
        CK(cudaMemcpyAsync(s.h_partial, s.scratch, 224, cudaMemcpyDeviceToHost, s.stream));
        CK(cudaStreamSynchronize(s.stream));
        CK(cudaGetLastError());
        memset(ctx->art.S1, 0, 96); memset(ctx->art.S2, 0, 96); memset(ctx->art.S3, 0, 96);
        memcpy(ctx->art.A, s.h_partial, 96); memcpy(ctx->art.B, s.h_partial + 96, 96);
        memcpy(ctx->art.sum_ry, s.h_partial + 192, 32);
    }
    *out = ctx->art;
    return KZGB_OK;
}

kzgb_ret kzgb_synth_instance(kzgb_ctx* ctx, uint64_t seed, uint64_t offset, size_t n, uint8_t* C, uint8_t* z, uint8_t* y,
                             uint8_t* pi, int out_on_device) {
    KZ_API_GUARD;
    if (!ctx || !C || !z || !y || !pi) return KZGB_BADARGS;
    DeviceSlot& s = ctx->slots[0];
    CK(cudaSetDevice(s.device));
    if (!s.comb_built) {
        CK(dmalloc(s.comb, 2 * 32 * 255));
        launch_build_comb(s.stream, s.comb);
        s.comb_built = true;
    }
    if (out_on_device) {
        launch_synth(s.stream, seed, offset, n, s.comb, C, z, y, pi);
        CK(cudaStreamSynchronize(s.stream));
    } else {
        size_t done = 0;
        while (done < n) {
            size_t k = n - done < s.n_max ? n - done : s.n_max;
            launch_synth(s.stream, seed, offset + done, k, s.comb, s.dC, s.dz, s.dy, s.dpi);
            CK(cudaMemcpyAsync(C + 48 * done, s.dC, 48 * k, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(z + 32 * done, s.dz, 32 * k, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(y + 32 * done, s.dy, 32 * k, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaMemcpyAsync(pi + 48 * done, s.dpi, 48 * k, cudaMemcpyDeviceToHost, s.stream));
            CK(cudaStreamSynchronize(s.stream));
            done += k;
        }
    }
    CK(cudaGetLastError());
    return KZGB_OK;
}
kzgb_ret kzgb_debug_op(kzgb_ctx* ctx, int op, const uint8_t* in, uint8_t* out, size_t count) {
    KZ_API_GUARD;
    if (!ctx || !in || !out) return KZGB_BADARGS;
    static const int isz[] = {0, 96, 48, 96, 96, 48, 48, 64, 64, 192, 96, 128, 96, 1152, 576, 576, 576, 576, 192, 64};
    static const int osz[] = {0, 48, 48, 48, 48, 48, 48, 32, 32, 96, 96, 96, 96, 576, 576, 576, 576, 576, 576, 32};
    if (op < 1 || op >= 19) return KZGB_BADARGS;                 // SHA256_64 is exercised through the FS stage exports
    DeviceSlot& s = ctx->slots[0];
    CK(cudaSetDevice(s.device));
    uint8_t *d_in = nullptr, *d_out = nullptr;
    CK(cudaMalloc((void**)&d_in, (size_t)isz[op] * count + 16)); CK(cudaMalloc((void**)&d_out, (size_t)osz[op] * count + 16));
    CK(cudaMemcpyAsync(d_in, in, (size_t)isz[op] * count, cudaMemcpyHostToDevice, s.stream));
    if (op <= 12) launch_debug_op(s.stream, op, d_in, d_out, count);
    else for (size_t i = 0; i < count; ++i) launch_pairing_debug(s.stream, op, s.lines, d_in + (size_t)isz[op] * i, d_out + (size_t)osz[op] * i);
    CK(cudaMemcpyAsync(out, d_out, (size_t)osz[op] * count, cudaMemcpyDeviceToHost, s.stream));
    CK(cudaStreamSynchronize(s.stream));
    CK(cudaFree(d_in)); CK(cudaFree(d_out));
    CK(cudaGetLastError());
    return KZGB_OK;
}

static kzgb_ret imad_rate(DeviceSlot& s, int mode, double* rate, double* ms_out) {
    CK(cudaSetDevice(s.device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, s.device));
    int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2000;
    launch_imad_bench(s.stream, (uint32_t*)s.scratch, blocks, threads, 100, mode);      // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(s.ev[9], s.stream));
        launch_imad_bench(s.stream, (uint32_t*)s.scratch, blocks, threads, iters, mode);
        CK(cudaEventRecord(s.ev[10], s.stream));
        CK(cudaStreamSynchronize(s.stream));
        float ms = ev_ms(s.ev[9], s.ev[10]);
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    *rate = (double)blocks * threads * (double)iters * 128.0 / (best * 1e-3);
    if (ms_out) *ms_out = best;
    return KZGB_OK;
}
kzgb_ret kzgb_imad_peak(kzgb_ctx* ctx, double* imad_per_sec_out, double* ms_out) {
    KZ_API_GUARD;
    if (!ctx || !imad_per_sec_out) return KZGB_BADARGS;
    return imad_rate(ctx->slots[0], 0, imad_per_sec_out, ms_out);
}
kzgb_ret kzgb_imad32_peak(kzgb_ctx* ctx, double* imad_per_sec_out, double* ms_out) {
    KZ_API_GUARD;
    if (!ctx || !imad_per_sec_out) return KZGB_BADARGS;
    return imad_rate(ctx->slots[0], 1, imad_per_sec_out, ms_out);
}
kzgb_ret kzgb_last_stage_ms(kzgb_ctx* ctx, float ms_out[KZGB_N_STAGES]) {
    if (!ctx || !ms_out) return KZGB_BADARGS;
    memcpy(ms_out, ctx->art.stage_ms, sizeof ctx->art.stage_ms);
    return KZGB_OK;
}
uint64_t kzgb_launch_count(const kzgb_ctx*) { return g_kzgb_launches.load(); }
int kzgb_set_threads(kzgb_ctx*, int) { return 0; }
kzgb_ret kzgb_set_subgroup_batch_min(kzgb_ctx* ctx, size_t n_min) {
    if (!ctx) return KZGB_BADARGS;
    for (auto& s : ctx->slots) s.sg_min = n_min;
    return KZGB_OK;
}

}  // extern "C"
