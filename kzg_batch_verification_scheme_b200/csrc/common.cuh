// Common macros.  Device math lives in headers as KZ_HD functions so that the SAME source can be
// compiled (a) by nvcc for sm_100a -- the product -- and (b) by g++ with -DKZGB_EMU into a host-side
// emulation library used ONLY by the no-GPU unit tests (tests/emu) to exercise the kernel logic in CI.
// The emulation replaces nothing in the product: libkzgb200.so contains no host arithmetic path.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(KZGB_EMU)
#define KZ_HD inline
#define KZ_COLD inline
#define KZ_CONSTANT static const
#define KZ_UNROLL
#else
#define KZ_HD __device__ __forceinline__
#define KZ_COLD static __device__ __noinline__
#define KZ_CONSTANT static __device__ __constant__
#define KZ_UNROLL _Pragma("unroll")
#endif

typedef uint8_t u8;
typedef uint32_t u32;
typedef uint64_t u64;

#define KZ_FS_CHUNK 128u      /* leaves per Fiat-Shamir chunk digest = KZGB_CHUNK of include/kzgb200.h (checked in host.cu) */
