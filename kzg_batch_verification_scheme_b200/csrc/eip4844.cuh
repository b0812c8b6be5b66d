// EIP-4844 / c-kzg-4844 transcript mode: one challenge r from a flat SHA-256 over the whole
// batch, coefficients r^0 .. r^(n-1) (255-bit), instead of the tree transcript with 128-bit r_i.  Device bodies; the
// flat hash itself is serial by construction and runs on the host (SHA extensions), see host_eip4844_* in k_fs.cu.
#pragma once
#include "field.cuh"

#define KZ_EIP_POW_BITS 32                    // batch sizes below 2^32

// table[k] = r^(2^k) (Montgomery form), r = int_be(hash) mod r_BLS; r_out = r (canonical limbs)
KZ_HD void eip_power_table(Fr* table, Fr& r_out, const u8* hash_be) {
    Fr raw;
    fr_raw_from_be(raw, hash_be);
    raw = fr_reduce_raw(raw);
    r_out = raw;
    Fr m = fr_to_mont(raw);
    for (int k = 0; k < KZ_EIP_POW_BITS; ++k) { table[k] = m; m = fr_mul(m, m); }
}
// r^i as canonical limbs: the raw value 1 times Montgomery-form factors stays canonical
KZ_HD Fr eip_power(const Fr* table, u64 i) {
    Fr acc = fr_zero();
    acc.v[0] = 1;
    for (int k = 0; k < KZ_EIP_POW_BITS; ++k)
        if ((i >> k) & 1) acc = fr_mul(acc, table[k]);
    return acc;
}
