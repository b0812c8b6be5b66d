/* kzgb200_testing.h -- exports of the CPU ORACLE library only (oracle/libkzgb_oracle.so): test infrastructure.
 * The product library (libkzgb200.so) does not export these; nothing in the product path uses them.
 * tests/, tools/gen_*.py and bench.py's CPU-baseline legs are the only callers. */
#ifndef KZGB200_TESTING_H
#define KZGB200_TESTING_H
#include "kzgb200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* insecure test setup from the documented tau = SHA256("kzgb200/insecure-test-tau") mod r (g1: n1*48 B, g2: n2*96 B) */
kzgb_ret kzgb_synth_setup(uint8_t *g1_monomial, size_t n1, uint8_t *g2_monomial, size_t n2);
/* n proofs of real polynomials with ncoef coefficients (BASELINE.json config[0]) */
kzgb_ret kzgb_oracle_synth_instance_poly(uint64_t seed, size_t n, size_t ncoef, uint8_t *C, uint8_t *z, uint8_t *y,
                                         uint8_t *pi, int threads);
/* planted invalid proof (config[2]): index for (seed, n), and the corruption of proof j (stays a valid G1 point) */
uint64_t kzgb_oracle_plant_index(uint64_t seed, uint64_t n);
kzgb_ret kzgb_oracle_plant_invalid(uint8_t *pi, size_t j);
/* pairing-free verdict with the known test tau: 1 iff A + tau*B = O */
int kzgb_oracle_tau_shortcut(const uint8_t A[96], const uint8_t B[96]);
/* status byte of one compressed point by the slow definition ([r]P = O) */
int kzgb_oracle_g1_status_slow(const uint8_t in[48]);
/* blob / cell instance generators: see oracle/blobs.hpp, oracle/cells.hpp */
#ifdef __cplusplus
}
#endif
#endif
