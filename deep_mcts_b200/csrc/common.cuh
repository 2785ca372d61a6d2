// Shared host/device helpers for libdeepmcts_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/deepmcts_b200.h"

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned char u8;

#define DM_FULL_MASK 0xffffffffu

// ---- error plumbing (no exception crosses the C ABI) -----------------------------------
void dm_set_error(const std::string& msg);
extern unsigned long long g_dm_launches;

#define DM_CHECK_ARG(cond, msg)                 \
    do {                                        \
        if (!(cond)) {                          \
            dm_set_error(std::string(msg));     \
            return DM_ERR_INVALID;              \
        }                                       \
    } while (0)

#define DM_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess) {                                                          \
            dm_set_error(std::string(#call) + ": " + cudaGetErrorString(e__));            \
            return DM_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

#define DM_LAUNCH_CHECK()                                                                  \
    do {                                                                                   \
        ++g_dm_launches;                                                                   \
        cudaError_t e__ = cudaGetLastError();                                              \
        if (e__ != cudaSuccess) {                                                          \
            dm_set_error(std::string("kernel launch: ") + cudaGetErrorString(e__));       \
            return DM_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

// ---- Philox4x32-10 counter-based RNG ---------------------------------------------------
struct Philox {
    u32 key0, key1;
    u32 c0, c1, c2, c3;
    u32 out[4];
    int have;

    __device__ __forceinline__ void init(u64 seed, u32 stream_lo, u32 stream_hi, u32 ctr) {
        key0 = (u32)seed;
        key1 = (u32)(seed >> 32);
        c0 = ctr;
        c1 = 0;
        c2 = stream_lo;
        c3 = stream_hi;
        have = 0;
    }
    __device__ __forceinline__ void round(u32& k0, u32& k1) {
        const u32 M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        u32 hi0 = __umulhi(M0, out[0]), lo0 = M0 * out[0];
        u32 hi1 = __umulhi(M1, out[2]), lo1 = M1 * out[2];
        u32 n0 = hi1 ^ out[1] ^ k0, n1 = lo1, n2 = hi0 ^ out[3] ^ k1, n3 = lo0;
        out[0] = n0; out[1] = n1; out[2] = n2; out[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    __device__ __forceinline__ void refill() {
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
        u32 k0 = key0, k1 = key1;
#pragma unroll
        for (int i = 0; i < 10; ++i) round(k0, k1);
        if (++c0 == 0) ++c1;
        have = 4;
    }
    __device__ __forceinline__ u32 next() {
        if (have == 0) refill();
        return out[--have];
    }
    // uniform integer in [0, n)
    __device__ __forceinline__ u32 below(u32 n) { return __umulhi(next(), n); }
    // uniform double in (0, 1)
    __device__ __forceinline__ double uniform() {
        u32 a = next(), b = next();
        u64 bits = ((u64)a << 21) ^ (u64)(b >> 11);  // 53 bits
        return ((double)(bits & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
    }
};
