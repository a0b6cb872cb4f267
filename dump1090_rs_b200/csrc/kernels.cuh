// dump1090_rs_b200/csrc/kernels.cuh -- device side of libb200adsb (sm_100a).
//
// The reference scans one sample at a time on one CPU thread
// (src/demod_2400.rs:115-212).  Here the same arithmetic is re-organised for a GPU.
// This file holds what the stage-1 kernel (scan7_kernel, scan7.cuh) shares with the rest
// (exact fast magnitude, CRC-24 by fields, classification, event table, gates), stage 2
// and the API-parity kernels.
//
//   stage 1 (scan7.cuh, one thread block per tile of T output positions)
//     IQ -> u16 magnitude (src/utils.rs:43-58); edge bits and the five PPM correlator signs of
//     every sample as bit planes de-interleaved modulo 12 samples (one Mode-S bit period at
//     2.4 Msps is 12/5 samples, so message bit n of try-phase t lives at 1/5-sample position
//     P = 5(j+19)+t+12n: sample P/5, correlator P%5)      (src/demod_2400.rs:62-83,158-182);
//     the five preamble templates as AND/shift of the edge planes, SNR + quiet-zone gates
//     (src/demod_2400.rs:127-146,215-321); per surviving position and try-phase five 23-bit
//     field extracts, DF, CRC-24 syndrome by table (GF(2)-linear), stateless classification
//     -> one 24-byte record; ICAO add-events by atomicMin
//                                          (src/mode_s/mod.rs:34-139, src/crc.rs:263-282)
//   finalize / resolve / emit / commit kernels
//         the sequential ICAO filter (src/icao_filter.rs) evaluated order-free:
//         member(a) at ordinal o  <=>  a == 0 || a in filter before the batch ||
//         firstAdd(a) < o; best-of-5 with the reference's strict '>' rule;
//         frames written in (buffer, j) order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200adsb.h"

namespace b200 {

constexpr int kTrailing = B200ADSB_TRAILING_SAMPLES;          // lib.rs:24
constexpr int kMaxSamples = B200ADSB_MODES_MAG_BUF_SAMPLES;    // lib.rs:22
constexpr int kMagLen = B200ADSB_MAG_DATA_LEN;
constexpr int kHaloFront = 2;    // tile mag index 0 <-> data index tile_start-2 (16 B aligned IQ loads)
constexpr int kHaloTot = 296;    // mags needed per tile = T + 296 (max tap j+289, +2 front, padded to 8)
constexpr int kLutWords = 12 * 25;          // field-extraction table: (residue of j+19, try_phase, field)
constexpr int kMaxTile = 8184;   // tile mag indices (< T+2) fit 13 bits; surv words <= 256
constexpr int kDefaultTile = 7384;   // 40 x 192 - 296: the halo-extended tile is 40 groups of 192 samples
constexpr int kTabWords = 256 + 256 + 64 + 256 + 8;   // CRC-24 field tables (see build_crc_tabs)
constexpr int kTab56 = 576;

// record word: kind<<29 | key (24 bit, or 1 for the "None" marker)
enum : uint32_t {
    K_NONE = 0, K_PAR_SHORT = 1, K_DF11_IID0 = 2, K_DF11_IID = 3, K_DF17 = 4, K_DF18 = 5,
    K_PAR_LONG = 6
};
constexpr uint32_t kNoneMarker = 1;  // kind NONE, key 1: score_modes_message returned None

// counters block (device, u32)
enum { C_POOL = 0, C_FLAGS = 1, C_EV_USED = 2, C_FRAMES = 3, C_CAND = 4, C_MEMBERS = 5,
       C_ADMIT = 6, C_NEWCNT = 7, C_TICKET = 8,
       C_STICKY = 9,    // enqueue-only batches: an earlier queued batch failed and the host has not acknowledged it
       C_TAILCUR = 10,  // carry mode: which of the two stream tails is current (flipped by a committed batch)
       C_WORDS = 16 };
// F_POOL_OVF .. F_REMOTE_BAD make a batch "bad": it is not committed to the filter (and, in carry mode,
// does not advance the stream tail).  F_SKIPPED / F_REMOTE_BAD also appear in d_result[1] of the
// enqueue-only calls (include/b200adsb.h).
enum : uint32_t { F_POOL_OVF = 1, F_EV_OVF = 2, F_FILTER_FULL = 4, F_SKIPPED = 8, F_REMOTE_BAD = 16 };
constexpr uint32_t kBadMask = F_POOL_OVF | F_EV_OVF | F_REMOTE_BAD;
constexpr int kTailWords = 352;   // kTrailing rounded up

constexpr uint32_t kMemberSlots = 8192;   // open addressing, >= 2 x 4096 keys
constexpr unsigned long long kNever = ~0ull;

struct ScanParams {
    const void *in;            // int16 (re,im) pairs, or u16 MagnitudeBuffer.data: buffer 0 of the BATCH
    const uint32_t *lengths;   // nullable per-buffer sample counts (indexed by batch buffer)
    uint32_t b0;               // first batch buffer of this launch (the host API scans per H2D chunk)
    uint32_t n_buffers;        // buffers of this launch
    uint32_t spb;              // samples per buffer (length when lengths == NULL)
    unsigned long long stride; // IQ: samples between buffers; MAG: u16 elements
    int T, tiles_per_buffer;
    int vec_ok;                // 16-byte aligned base and stride % 4 == 0
    uint32_t *rec;             // pool of 6-word records {j, w[5]}
    uint32_t pool_cap;
    uint2 *tile_dir;           // per tile of the batch: (pool base, count)
    uint32_t *counters;
    uint32_t *ev_keys;
    unsigned long long *ev_ord;
    uint32_t *ev_used;
    uint32_t ev_mask;
    unsigned long long ord_first, ord_stride;
    const uint32_t *crc_lanes; // [kLaneTabs][32]: CRC-24 field tables, one entry per lane (a112_sh / a56_sh)
    const uint32_t *lut;       // [12][5][5] field extraction table for this tile size
    // opt-in stream continuity (B200ADSB_OPT_CARRY): the 326 leading MagnitudeBuffer slots of a
    // buffer hold the previous buffer's last samples instead of zeros
    int carry;
    const uint32_t *tails;     // two stream tails of kTailWords each; counters[C_TAILCUR] selects the current one
};

// ------------------------------------------------------------------ magnitude
// src/utils.rs:47-55: fi = im/2^15, fq = re/2^15 (exact), fma(fi,fi, rn(fq*fq)),
// IEEE sqrt, fma(mag, 65535, 0.5), saturating truncation (Rust `as u16`).
// Reference form (IEEE intrinsics), used by to_mag_kernel/emit and as the in-library
// check of the fast form below (b200adsb_debug_mag_sweep compares all 2^32 inputs).
__device__ __forceinline__ uint32_t mag_u16(int re, int im)
{
    const float fi = __fmul_rn(__int2float_rn(im), 0x1p-15f);
    const float fq = __fmul_rn(__int2float_rn(re), 0x1p-15f);
    const float q2 = __fmul_rn(fq, fq);
    const float msq = __fmaf_rn(fi, fi, q2);
    const float mag = __fsqrt_rn(msq);
    const float v = fminf(__fmaf_rn(mag, 65535.0f, 0.5f), 65535.0f);
    return __float2uint_rz(v);
}
__device__ __forceinline__ uint32_t mag_pair(uint32_t w)  // w = re | im<<16
{
    return mag_u16((int)(short)(w & 0xffffu), (int)(short)(w >> 16));
}

// Fast form: two samples per instruction with Blackwell's packed f32x2 pipe (FFMA2/FMUL2/
// FADD2), bit-identical to the form above on every (re, im) in int16^2 (exhaustively
// checked on the GPU by b200adsb_debug_mag_sweep):
//  * int16 -> f32 without I2F: bytes of (v ^ 0x8000) under exponent 0x4B give
//    2^23 + v + 32768, and fma(f, 2^-15, -257) = v / 2^15 exactly;
//  * sqrt: MUFU.RSQ seed y, s0 = x*y, then s = s0 + (x - s0^2) * y/2 with the residual
//    in FMA form, correctly rounded on the reachable set {0} U [2^-30, 2] (x + 1e-30 == x
//    there, and keeps rsqrt finite at 0); the /2 is folded as exact power-of-two scalings;
//  * saturating truncation: min(v, 65535) then add 2^23 rounding toward zero, so the
//    result's low mantissa bits ARE the integer.
// Returns the f32 bit patterns 0x4B000000 + magnitude (differences of these are
// differences of magnitudes; the low 16 bits are the u16).
typedef unsigned long long u64x;
__device__ __forceinline__ u64x f2_pack(float a, float b)
{
    u64x r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void f2_unpack(u64x v, float &a, float &b)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ u64x f2_fma(u64x a, u64x b, u64x c)
{
    u64x r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64x f2_mul(u64x a, u64x b)
{
    u64x r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64x f2_add(u64x a, u64x b)
{
    u64x r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64x f2_sub(u64x a, u64x b)
{
    u64x r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64x f2_add_rz(u64x a, u64x b)
{
    u64x r;
    asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
#ifndef B200_MAG_V2
#define B200_MAG_V2 1
#endif
#if B200_MAG_V2
// v2: I2F.S16 straight from the halves of the IQ word (no XOR/PRMT/rescale), everything kept
// scaled by 2^15 (powers of two commute with every rounding here, nothing over/underflows:
// x' = x * 2^30 <= 2^31), residual with the negated operand folded into FFMA2:
//   s0 = x*y, e = fma(-s0, s0, x), s = fma(e, y/2, s0)   (same reals, same roundings as v1)
__device__ __forceinline__ float cvt_s16_lo(uint32_t w)
{
    float f;
    asm("{ .reg .b16 l, h; mov.b32 {l, h}, %1; cvt.rn.f32.s16 %0, l; }" : "=f"(f) : "r"(w));
    return f;
}
__device__ __forceinline__ float cvt_s16_hi(uint32_t w)
{
    float f;
    asm("{ .reg .b16 l, h; mov.b32 {l, h}, %1; cvt.rn.f32.s16 %0, h; }" : "=f"(f) : "r"(w));
    return f;
}
__device__ __forceinline__ u64x mag_pair_fast2(uint32_t wa, uint32_t wb)
{
    const u64x fq = f2_pack(cvt_s16_lo(wa), cvt_s16_lo(wb));
    const u64x fi = f2_pack(cvt_s16_hi(wa), cvt_s16_hi(wb));
    const u64x x = f2_fma(fi, fi, f2_mul(fq, fq));
    float xa, xb, ya, yb;
    f2_unpack(f2_add(x, f2_pack(1e-20f, 1e-20f)), xa, xb);
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ya) : "f"(xa));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yb) : "f"(xb));
    const u64x y = f2_pack(ya, yb);
    const u64x s0 = f2_mul(x, y);
    const u64x yh = f2_mul(y, f2_pack(0.5f, 0.5f));
    float s0a, s0b;
    f2_unpack(s0, s0a, s0b);
    const u64x e = f2_fma(f2_pack(-s0a, -s0b), s0, x);          // x - s0^2, one rounding
    const u64x s = f2_fma(e, yh, s0);
    float va, vb;
    f2_unpack(f2_fma(s, f2_pack(0x1.fffep0f, 0x1.fffep0f), f2_pack(0.5f, 0.5f)), va, vb);   // 65535 * 2^-15
    return f2_add_rz(f2_pack(fminf(va, 65535.0f), fminf(vb, 65535.0f)), f2_pack(8388608.0f, 8388608.0f));
}
#else
__device__ __forceinline__ u64x mag_pair_fast2(uint32_t wa, uint32_t wb)
{
    const uint32_t ta = wa ^ 0x80008000u, tb = wb ^ 0x80008000u;
    const u64x fre = f2_pack(__uint_as_float(__byte_perm(ta, 0x4B000000u, 0x7610)),
                             __uint_as_float(__byte_perm(tb, 0x4B000000u, 0x7610)));
    const u64x fim = f2_pack(__uint_as_float(__byte_perm(ta, 0x4B000000u, 0x7632)),
                             __uint_as_float(__byte_perm(tb, 0x4B000000u, 0x7632)));
    const u64x c15 = f2_pack(0x1p-15f, 0x1p-15f), m257 = f2_pack(-257.0f, -257.0f);
    const u64x fq = f2_fma(fre, c15, m257), fi = f2_fma(fim, c15, m257);
    const u64x x = f2_fma(fi, fi, f2_mul(fq, fq));
    float xa, xb, ya, yb;
    f2_unpack(f2_add(x, f2_pack(1e-30f, 1e-30f)), xa, xb);
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(ya) : "f"(xa));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yb) : "f"(xb));
    const u64x y = f2_pack(ya, yb);
    const u64x s0 = f2_mul(x, y);
    const u64x t = f2_mul(s0, f2_pack(-0.5f, -0.5f));          // -s0/2 (exact)
    const u64x hx = f2_mul(x, f2_pack(0.5f, 0.5f));            // x/2 (exact)
    const u64x eh = f2_fma(t, s0, hx);                          // (x - s0^2)/2, one rounding
    const u64x s = f2_fma(eh, y, s0);                           // s0 + (x - s0^2) * y/2
    float va, vb;
    f2_unpack(f2_fma(s, f2_pack(65535.0f, 65535.0f), f2_pack(0.5f, 0.5f)), va, vb);
    return f2_add_rz(f2_pack(fminf(va, 65535.0f), fminf(vb, 65535.0f)), f2_pack(8388608.0f, 8388608.0f));
}
#endif
__device__ __forceinline__ uint32_t mag_bits_fast(uint32_t w)
{
    float a, b;
    f2_unpack(mag_pair_fast2(w, w), a, b);
    return __float_as_uint(a);
}

// ------------------------------------------------------------------ CRC-24 by fields
// Message bit n (MSB first) = 5m + r lives in field r, bit m.  syndrome =
// sum_n b_n x^(L-1-n) mod G  (src/crc.rs:263-282 is exactly M(x) mod G).
__device__ __forceinline__ uint32_t mulx(uint32_t s)
{
    s <<= 1;
    return (s & 0x1000000u) ? (s ^ 0x1FFF409u) : s;
}
// Field sums a112(f) = sum_{m<=21} f[m] x^(107-5m), a56(f) = sum_{m<=10} f[m] x^(51-5m) with the tables held
// in registers, one entry per lane, and looked up with warp shuffles (5-bit chunks of the field: shfl takes
// the lane index modulo 32, so `f >> 5c` needs no mask).  A table lookup through L1 costs ~7 tag wavefronts
// per warp on the LSU pipe, a shuffle one, and 70 instead of 103 instructions per long message
// (profiles/r2).  All 32 lanes must execute these together.
//   lanes[c][v], c = 0..4: sum_{b<5, 5c+b<=21} v[b] x^(107-5(5c+b))      (long)
//   lanes[5+c][v], c = 0..1: sum_{b<5} v[b] x^(51-5(5c+b))               (short; m = 10 is x^1 = 2)
constexpr int kLaneTabs = 7;
struct CrcLanes {
    uint32_t t[kLaneTabs];
};
__device__ __forceinline__ uint32_t a112_sh(const CrcLanes &L, uint32_t f)
{
    return __shfl_sync(0xffffffffu, L.t[0], f) ^ __shfl_sync(0xffffffffu, L.t[1], f >> 5) ^
           __shfl_sync(0xffffffffu, L.t[2], f >> 10) ^ __shfl_sync(0xffffffffu, L.t[3], f >> 15) ^
           __shfl_sync(0xffffffffu, L.t[4], (f >> 20) & 3u);
}
__device__ __forceinline__ uint32_t a56_sh(const CrcLanes &L, uint32_t f)
{
    return __shfl_sync(0xffffffffu, L.t[5], f) ^ __shfl_sync(0xffffffffu, L.t[6], f >> 5) ^ (((f >> 10) & 1u) << 1);
}
__device__ __forceinline__ uint32_t syn112_fields_sh(const CrcLanes &L, const uint32_t f[5])
{
    uint32_t s = a112_sh(L, f[0]);
    s = mulx(s) ^ a112_sh(L, f[1]);
    s = mulx(s) ^ a112_sh(L, f[2]);
    s = mulx(s) ^ a112_sh(L, f[3]);
    s = mulx(s) ^ a112_sh(L, f[4]);
    return s ^ (((f[0] >> 22) & 1u) << 1) ^ ((f[1] >> 22) & 1u);   // bits 110, 111
}
__device__ __forceinline__ uint32_t syn56_fields_sh(const CrcLanes &L, const uint32_t f[5])
{
    uint32_t s = a56_sh(L, f[0]);
    s = mulx(s) ^ a56_sh(L, f[1]);
    s = mulx(s) ^ a56_sh(L, f[2]);
    s = mulx(s) ^ a56_sh(L, f[3]);
    s = mulx(s) ^ a56_sh(L, f[4]);
    return s ^ ((f[0] >> 11) & 1u);                                 // bit 55
}
// bits n0 .. n0+cnt-1 of the message, MSB first, from the five fields
template <int N0, int CNT>
__device__ __forceinline__ uint32_t msg_bits(const uint32_t f[5])
{
    uint32_t v = 0;
#pragma unroll
    for (int n = N0; n < N0 + CNT; n++)
        v = (v << 1) | ((f[n % 5] >> (n / 5)) & 1u);
    return v;
}

// byte-wise form for the message-level entry points (crc.rs:263-282 verbatim in spirit)
__device__ __forceinline__ uint32_t crc_bytes(const uint32_t *tab256, const uint8_t *m, int nbytes)
{
    uint32_t rem = 0;
    for (int i = 0; i < nbytes - 3; i++)
        rem = ((rem << 8) ^ tab256[m[i] ^ ((rem >> 16) & 0xff)]) & 0xffffffu;
    return rem ^ ((uint32_t)m[nbytes - 3] << 16) ^ ((uint32_t)m[nbytes - 2] << 8) ^ m[nbytes - 1];
}

__device__ __forceinline__ uint32_t classify_bytes(const uint32_t *tab256, const uint8_t *m)
{
    uint32_t any = 0;
    for (int i = 0; i < 14; i++)
        any |= m[i];
    if (!any)
        return kNoneMarker;
    const uint32_t df = m[0] >> 3, bit = 1u << df;
    const uint32_t addr = ((uint32_t)m[1] << 16) | ((uint32_t)m[2] << 8) | m[3];
    if (bit & 0x00000031u)
        return (K_PAR_SHORT << 29) | crc_bytes(tab256, m, 7);
    if (df == 11) {
        const uint32_t syn = crc_bytes(tab256, m, 7);
        if (syn & 0xffff80u)
            return 0;
        return (((syn & 0x7f) ? K_DF11_IID : K_DF11_IID0) << 29) | addr;
    }
    if (bit & 0x00060000u) {
        if (crc_bytes(tab256, m, 14) != 0)
            return 0;
        return ((df == 17 ? K_DF17 : K_DF18) << 29) | addr;
    }
    if (bit & 0xFF310000u)
        return (K_PAR_LONG << 29) | crc_bytes(tab256, m, 14);
    return 0;
}

// ------------------------------------------------------------------ event table
__device__ __forceinline__ uint32_t hash32(uint32_t k) { return (k * 2654435761u) >> 7; }

// firstAdd(key) = min(firstAdd(key), ord)   
__device__ inline void event_add(uint32_t *ev_keys, unsigned long long *ev_ord, uint32_t *ev_used,
                                 uint32_t mask, uint32_t *counters, uint32_t key,
                                 unsigned long long ord)
{
    if ((key & 0xffffffu) == 0)
        return;   // address 0 always tests true (icao_filter.rs:71,78): neither DF17/11 nor DF18 add it
    uint32_t h = hash32(key) & mask;
    for (uint32_t probe = 0; probe <= mask; probe++) {
        const uint32_t prev = atomicCAS(&ev_keys[h], 0u, key);
        if (prev == 0u) {
            const uint32_t u = atomicAdd(&counters[C_EV_USED], 1u);
            ev_used[u] = h;   // u <= mask because every slot is claimed once
        }
        if (prev == 0u || prev == key) {
            atomicMin(&ev_ord[h], ord);
            return;
        }
        h = (h + 1) & mask;
    }
    atomicOr(&counters[C_FLAGS], F_EV_OVF);
}

// 64 Kbit membership pre-filter over (filter keys U this batch's event keys): one load rejects
// the overwhelming majority of address-parity syndromes, which are not aircraft addresses
constexpr uint32_t kBloomWords = 2048;
__device__ __forceinline__ uint32_t bloom_bit(uint32_t key) { return (key * 0x9E3779B1u) >> 16; }
__device__ __forceinline__ bool bloom_hit(const uint32_t *bloom, uint32_t key)
{
    const uint32_t h = bloom_bit(key);
    return (bloom[h >> 5] >> (h & 31)) & 1u;
}

__device__ __forceinline__ bool members_has(const uint32_t *members, uint32_t key)
{
    uint32_t h = hash32(key) & (kMemberSlots - 1);
    for (;;) {
        const uint32_t v = members[h];
        if (v == key)
            return true;
        if (v == 0u)
            return false;
        h = (h + 1) & (kMemberSlots - 1);
    }
}
__device__ __forceinline__ void members_insert(uint32_t *members, uint32_t key)
{
    uint32_t h = hash32(key) & (kMemberSlots - 1);
    for (;;) {
        const uint32_t prev = atomicCAS(&members[h], 0u, key);
        if (prev == 0u || prev == key)
            return;
        h = (h + 1) & (kMemberSlots - 1);
    }
}
__device__ __forceinline__ unsigned long long event_first(const uint32_t *ev_keys,
                                                          const unsigned long long *ev_ord,
                                                          uint32_t mask, uint32_t key)
{
    uint32_t h = hash32(key) & mask;
    for (uint32_t probe = 0; probe <= mask; probe++) {
        const uint32_t v = ev_keys[h];
        if (v == key)
            return ev_ord[h];
        if (v == 0u)
            return kNever;
        h = (h + 1) & mask;
    }
    return kNever;
}

// ================================================================== scan kernel
// dd layout: first differences d[i] = m[i+1]-m[i] of the tile's magnitudes, 384 per block
// padded to 396 words; the 12 pad words repeat the next block's first 12, so a lane that
// walks one residue class (i = 12q + rho, q = 32 consecutive) reads d[i..i+2] with plain
// strides and 32 consecutive (block, rho) items hit 32 different banks.

// where the samples before a buffer come from in carry mode (B200ADSB_OPT_CARRY): the previous
// buffers of the batch, then the stream tail saved by the previous batch
struct CarrySrc {
    const uint32_t *in;        // whole batch (buffer 0 of the batch, not of the launch)
    unsigned long long stride;
    const uint32_t *lengths;
    uint32_t spb;
    const uint32_t *tails;     // the two saved stream tails (last kTrailing samples before this batch)
    const uint32_t *counters;  // [C_TAILCUR] selects the current tail
    int on;
};
// IQ word of sample s of batch buffer b; s < 0 reaches back into the stream (carry mode) or is
// zero (magnitude(0, 0) = 0 is exactly the MagnitudeBuffer zero fill, lib.rs:36-44); s >= len is zero
__device__ __forceinline__ uint32_t iq_word(const uint32_t *b32, int s, int len, const CarrySrc &cs, uint32_t b)
{
    if (s >= 0)
        return s < len ? __ldg(b32 + s) : 0u;
    if (!cs.on)
        return 0u;
    while (b > 0) {            // buffers shorter than the reach-back are walked through
        b--;
        const int pl = cs.lengths ? (int)min(cs.lengths[b], cs.spb) : (int)cs.spb;
        s += pl;
        if (s >= 0)
            return __ldg(cs.in + (unsigned long long)b * cs.stride + (unsigned)s);
    }
    const int k = kTrailing + s;
    return k >= 0 ? cs.tails[kTailWords * (cs.counters[C_TAILCUR] & 1u) + k] : 0u;
}

// SNR and quiet-zone gates of one template match (demod_2400.rs:129,135-146) with the
// template's high/signal/noise (demod_2400.rs:226-317); cs = template case 0..4; pp = the 19
// magnitudes from the candidate position on, jl = its survivor bit.
__device__ __forceinline__ void gate_eval(const uint16_t *pp, uint32_t *surv, int jl, uint32_t cs)
{
    int high;
    uint32_t sig, noise;
    switch (cs) {
    case 0:
        high = ((int)pp[1] + pp[3] + pp[9] + pp[11] + pp[12]) / 4;
        sig = (uint32_t)pp[1] + pp[3] + pp[9];
        noise = (uint32_t)pp[5] + pp[6] + pp[7];
        break;
    case 1:
        high = ((int)pp[1] + pp[3] + pp[9] + pp[12]) / 4;
        sig = (uint32_t)pp[1] + pp[3] + pp[9] + pp[12];
        noise = (uint32_t)pp[5] + pp[6] + pp[7] + pp[8];
        break;
    case 2:
        high = ((int)pp[1] + pp[3] + pp[4] + pp[9] + pp[10] + pp[12]) / 4;
        sig = (uint32_t)pp[1] + pp[12];
        noise = (uint32_t)pp[6] + pp[7];
        break;
    case 3:
        high = ((int)pp[1] + pp[4] + pp[10] + pp[12]) / 4;
        sig = (uint32_t)pp[1] + pp[4] + pp[10] + pp[12];
        noise = (uint32_t)pp[5] + pp[6] + pp[7] + pp[8];
        break;
    default:
        high = ((int)pp[1] + pp[2] + pp[4] + pp[10] + pp[12]) / 4;
        sig = (uint32_t)pp[4] + pp[10] + pp[12];
        noise = (uint32_t)pp[6] + pp[7] + pp[8];
        break;
    }
    if (sig * 2 < 3 * noise)   // demod_2400.rs:129
        return;
    const int mx = max(max(max((int)pp[5], (int)pp[6]), max((int)pp[7], (int)pp[8])),
                       max(max(max((int)pp[14], (int)pp[15]), max((int)pp[16], (int)pp[17])), (int)pp[18]));
    if (mx >= high)            // demod_2400.rs:135-146
        return;
    atomicOr(&surv[jl >> 5], 1u << (jl & 31));
}

__device__ __forceinline__ uint32_t df_of_fields(const uint32_t f[5])
{
    return ((f[0] & 1u) << 4) | ((f[1] & 1u) << 3) | ((f[2] & 1u) << 2) | ((f[3] & 1u) << 1) | (f[4] & 1u);
}

// ================================================================== to_mag kernel
// utils::to_mag (src/utils.rs:43-58) with the MagnitudeBuffer layout of src/lib.rs:29-51
__global__ void to_mag_kernel(const uint32_t *__restrict__ iq, int n, uint16_t *__restrict__ data)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kMagLen)
        return;
    const int s = i - kTrailing;
    data[i] = (s >= 0 && s < n) ? (uint16_t)mag_pair(__ldg(iq + s)) : (uint16_t)0;
}

// exhaustive check of mag_bits_fast against the IEEE form over all 2^32 (re, im) inputs
__global__ void mag_sweep_kernel(unsigned long long *mismatches, uint32_t *first_bad)
{
    const uint32_t hi = blockIdx.x * blockDim.x + threadIdx.x;   // 2^24 threads x 256 inputs
    uint32_t bad = 0;
    for (uint32_t lo = 0; lo < 256; lo++) {
        const uint32_t w = (hi << 8) | lo;
        const uint32_t fast = mag_bits_fast(w);
        if ((fast & 0xffffu) != mag_pair(w) || (fast >> 16) != 0x4B00u) {
            bad++;
            atomicMin(first_bad, w);
        }
    }
    if (bad)
        atomicAdd(mismatches, (unsigned long long)bad);
}

// CU8 ingest: 8-bit unsigned (I, Q) pairs -> CS16 with the conversion SoapySDR's RTL-SDR module applies when
// asked for CS16 (the reference's default driver, dump1090_rs/src/main.rs:49-55,143):
//   int16((float(u8) - 127.4f) * (1.0f / 128.0f) * 32767.0f), truncating.  Two samples per thread.
__device__ __forceinline__ uint32_t cu8_to_cs16(uint32_t v)
{
    const float f = __fmul_rn(__fmul_rn(__fsub_rn(__uint2float_rn(v), 127.4f), 1.0f / 128.0f), 32767.0f);
    return (uint32_t)__float2int_rz(f) & 0xffffu;
}
__global__ void cu8_expand_kernel(const uint32_t *__restrict__ in, uint2 *__restrict__ out, size_t words)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= words)
        return;
    const uint32_t w = __ldg(in + i);       // I0 Q0 I1 Q1
    out[i] = make_uint2(cu8_to_cs16(w & 0xffu) | (cu8_to_cs16((w >> 8) & 0xffu) << 16),
                        cu8_to_cs16((w >> 16) & 0xffu) | (cu8_to_cs16(w >> 24) << 16));
}

// ================================================================== message-level kernels
__global__ void checksum_kernel(const uint8_t *__restrict__ msgs, int n, int nbytes,
                                const uint32_t *__restrict__ tab256, uint32_t *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = crc_bytes(tab256, msgs + 14ll * i, nbytes);
}

// one record per message (w[0] only), tile_dir describing "tiles" of 1024 messages
__global__ void classify_msgs_kernel(const uint8_t *__restrict__ msgs, int n,
                                     const uint32_t *__restrict__ tab256, uint32_t *rec,
                                     uint2 *tile_dir, int per_tile, uint32_t *counters,
                                     uint32_t *ev_keys, unsigned long long *ev_ord, uint32_t *ev_used,
                                     uint32_t ev_mask, unsigned long long ord_first)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t wd = classify_bytes(tab256, msgs + 14ll * i);
        uint32_t *r = rec + 6ull * i;
        r[0] = (uint32_t)(i % per_tile);
        r[1] = wd;
        r[2] = r[3] = r[4] = r[5] = 0;
        const uint32_t kind = wd >> 29;
        if (kind == K_DF11_IID0 || kind == K_DF17 || kind == K_DF18) {
            const uint32_t key = (wd & 0xffffffu) | (kind == K_DF18 ? B200ADSB_ICAO_FILTER_ADSB_NT : 0u);
            const unsigned long long ord =
                ((ord_first + (unsigned long long)(i / per_tile)) << 20) |
                ((unsigned long long)(i % per_tile) << 3);
            event_add(ev_keys, ev_ord, ev_used, ev_mask, counters, key, ord);
        }
        if (i % per_tile == 0)
            tile_dir[i / per_tile] = make_uint2((uint32_t)i, (uint32_t)min(per_tile, n - i));
    }
}

// ================================================================== events
// export (key, ord) pairs of the slots used by this rank
__global__ void events_export_kernel(const uint32_t *ev_keys, const unsigned long long *ev_ord,
                                     const uint32_t *ev_used, const uint32_t *counters,
                                     unsigned long long *pairs, uint32_t cap)
{
    const uint32_t n = min(counters[C_EV_USED], cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t h = ev_used[i];
        pairs[2ull * i] = ev_keys[h];
        pairs[2ull * i + 1] = ev_ord[h];
    }
}
// packed form for a sync-free exchange: rows[0] = (count, this rank's bad-batch flags),
// rows[1..] = (key, ordinal).  A rank with more events than the exchange buffer holds fails its own
// batch too (F_EV_OVF), so that every rank takes the same decision.
__global__ void events_pack_kernel(const uint32_t *ev_keys, const unsigned long long *ev_ord,
                                   const uint32_t *ev_used, uint32_t *counters,
                                   unsigned long long *rows, uint32_t rows_cap)
{
    const uint32_t n = counters[C_EV_USED];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        uint32_t fl = counters[C_FLAGS] & kBadMask;
        if (n > rows_cap - 1) {
            fl |= F_EV_OVF;
            atomicOr(&counters[C_FLAGS], F_EV_OVF);
        }
        rows[0] = n;
        rows[1] = fl;
    }
    const uint32_t m = min(n, rows_cap - 1);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const uint32_t h = ev_used[i];
        rows[2ull * (i + 1)] = ev_keys[h];
        rows[2ull * (i + 1) + 1] = ev_ord[h];
    }
}
// gathered: n_ranks blocks of rows_per_rank packed rows; merges every block but `skip` (this rank's
// own); a bad batch on any rank makes this rank's batch bad too (F_REMOTE_BAD)
__global__ void events_import_packed_kernel(const unsigned long long *gathered, uint32_t n_ranks,
                                            uint32_t rows_per_rank, uint32_t skip, uint32_t *ev_keys,
                                            unsigned long long *ev_ord, uint32_t *ev_used, uint32_t mask,
                                            uint32_t *counters)
{
    for (uint32_t r = 0; r < n_ranks; r++) {
        if (r == skip)
            continue;
        const unsigned long long *rows = gathered + 2ull * r * rows_per_rank;
        const unsigned long long n = rows[0];
        if (n > rows_per_rank - 1 || (rows[1] & kBadMask) != 0ull) {
            if (blockIdx.x == 0 && threadIdx.x == 0)
                atomicOr(&counters[C_FLAGS], F_REMOTE_BAD);
            continue;
        }
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)n; i += gridDim.x * blockDim.x)
            event_add(ev_keys, ev_ord, ev_used, mask, counters, (uint32_t)rows[2ull * (i + 1)], rows[2ull * (i + 1) + 1]);
    }
}

// ---- the same exchange fused with its transport: no NCCL call between scan and resolve.
// Every rank owns a buffer in symmetric (peer-mapped) memory, all with the same layout:
//   u64 flags[2][world]                      flag [p][r] = epoch of the last block rank r pushed with parity p
//   u64 blocks[2][world][rows_per_rank][2]   packed rows as above
// events_push_symm_kernel packs this rank's events and stores them straight into block [p][rank] of EVERY
// rank's buffer over NVLink (peer stores), fences, and the last block to finish raises flag [p][rank] on
// every rank.  events_import_symm_kernel (next launch on the same stream) waits until all `world` flags of
// parity p have reached the epoch, then merges the blocks found locally.  Two parities: a rank can be at
// most one exchange ahead of a peer (its next push needs the peer's push of this epoch), so the block a
// slow peer is still reading is never the one being overwritten.
__device__ __forceinline__ unsigned long long *symm_block(unsigned long long *base, uint32_t world, uint32_t rows,
                                                          uint32_t parity, uint32_t r)
{
    return base + 2ull * world + 2ull * rows * ((unsigned long long)parity * world + r);
}
__global__ void events_push_symm_kernel(const uint32_t *ev_keys, const unsigned long long *ev_ord,
                                        const uint32_t *ev_used, uint32_t *counters,
                                        unsigned long long *const *peer_bufs, uint32_t rank, uint32_t world,
                                        uint32_t rows_per_rank, unsigned long long epoch, uint32_t force_flags)
{
    __shared__ uint32_t s_last;
    const uint32_t parity = (uint32_t)(epoch & 1ull);
    const uint32_t n = counters[C_EV_USED];
    const uint32_t m = min(n, rows_per_rank - 1);
    if (blockIdx.x == 0 && threadIdx.x < world) {
        uint32_t fl = (counters[C_FLAGS] & kBadMask) | force_flags;
        if (n > rows_per_rank - 1) {
            fl |= F_EV_OVF;
            if (threadIdx.x == 0)
                atomicOr(&counters[C_FLAGS], F_EV_OVF);
        }
        unsigned long long *dst = symm_block(peer_bufs[threadIdx.x], world, rows_per_rank, parity, rank);
        dst[0] = n;
        dst[1] = fl;
    }
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const uint32_t h = ev_used[i];
        const unsigned long long key = ev_keys[h], ord = ev_ord[h];
        for (uint32_t q = 0; q < world; q++) {
            unsigned long long *dst = symm_block(peer_bufs[q], world, rows_per_rank, parity, rank);
            dst[2ull * (i + 1)] = key;
            dst[2ull * (i + 1) + 1] = ord;
        }
    }
    __threadfence_system();                    // this thread's peer stores before the ticket
    __syncthreads();
    if (threadIdx.x == 0)
        s_last = atomicAdd(&counters[C_TICKET], 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < world) {
            unsigned long long *flag = peer_bufs[threadIdx.x] + (unsigned long long)parity * world + rank;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
        }
        if (threadIdx.x == 0)
            counters[C_TICKET] = 0;
    }
}
__global__ void events_import_symm_kernel(unsigned long long *local_buf, uint32_t rank, uint32_t world,
                                          uint32_t rows_per_rank, unsigned long long epoch, uint32_t *ev_keys,
                                          unsigned long long *ev_ord, uint32_t *ev_used, uint32_t mask,
                                          uint32_t *counters)
{
    __shared__ uint32_t s_timeout;
    const uint32_t parity = (uint32_t)(epoch & 1ull);
    if (threadIdx.x == 0)
        s_timeout = 0;
    __syncthreads();
    if (threadIdx.x < world) {
        const unsigned long long *flag = local_buf + (unsigned long long)parity * world + threadIdx.x;
        const long long t0 = clock64();
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
            if (v >= epoch)
                break;
            if (clock64() - t0 > 60000000000ll) {  // ~30 s: a peer died; fail the batch instead of hanging
                s_timeout = 1;
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (s_timeout) {
        if (blockIdx.x == 0 && threadIdx.x == 0)
            atomicOr(&counters[C_FLAGS], F_REMOTE_BAD);
        return;
    }
    for (uint32_t r = 0; r < world; r++) {
        if (r == rank)
            continue;
        // (written by a peer over NVLink: read through L2, never from this SM's L1)
        const unsigned long long *rows = symm_block(local_buf, world, rows_per_rank, parity, r);
        const unsigned long long n = __ldcg(rows);
        if (n > rows_per_rank - 1 || (__ldcg(rows + 1) & kBadMask) != 0ull) {
            if (blockIdx.x == 0 && threadIdx.x == 0)
                atomicOr(&counters[C_FLAGS], F_REMOTE_BAD);
            continue;
        }
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)n; i += gridDim.x * blockDim.x)
            event_add(ev_keys, ev_ord, ev_used, mask, counters, (uint32_t)__ldcg(rows + 2ull * (i + 1)),
                      __ldcg(rows + 2ull * (i + 1) + 1));
    }
}

__global__ void events_import_kernel(const unsigned long long *pairs, uint32_t n, uint32_t *ev_keys,
                                     unsigned long long *ev_ord, uint32_t *ev_used, uint32_t mask,
                                     uint32_t *counters)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        event_add(ev_keys, ev_ord, ev_used, mask, counters, (uint32_t)pairs[2ull * i], pairs[2ull * i + 1]);
}

// Capacity rule of icao_filter_add (src/icao_filter.rs:46-62): the table holds the
// first 4096 distinct keys by first-add order; later ones are dropped.  One block.
struct FinalizeArgs {
    uint32_t *ev_keys;
    unsigned long long *ev_ord;
    const uint32_t *ev_used;
    uint32_t *ev_tmp, *new_keys, *counters, *members;
    uint32_t ev_mask;
    uint32_t *bloom;
};
__device__ __forceinline__ void events_finalize_body(
    const uint32_t *ev_keys, unsigned long long *ev_ord, const uint32_t *ev_used, uint32_t *ev_tmp,
    uint32_t *new_keys, uint32_t *counters, const uint32_t *members, uint32_t ev_mask, uint32_t *bloom)
{
    const uint32_t n = counters[C_EV_USED];
    const uint32_t have = counters[C_MEMBERS];
    for (uint32_t i = threadIdx.x; i < kBloomWords; i += blockDim.x)
        bloom[i] = 0u;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kMemberSlots; i += blockDim.x) {
        const uint32_t k = members[i];
        if (k) {
            const uint32_t h = bloom_bit(k);
            atomicOr(&bloom[h >> 5], 1u << (h & 31));
        }
    }
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t h = bloom_bit(ev_keys[ev_used[i]]);
        atomicOr(&bloom[h >> 5], 1u << (h & 31));
    }
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t h = ev_used[i];
        const uint32_t key = ev_keys[h];
        bool is_new = !members_has(members, key);
        if (is_new && (key & B200ADSB_ICAO_FILTER_ADSB_NT)) {
            // DF18 adds addr|ADSB_NT only when icao_filter_test(addr) was false at that
            // moment (mode_s/mod.rs:97-104): not if the plain address was in the filter
            // before the batch or was first added earlier in it.  (If that earlier add was
            // dropped because the table was full, this add is dropped for the same reason.)
            const uint32_t plain = key & 0xffffffu;
            if (members_has(members, plain) || event_first(ev_keys, ev_ord, ev_mask, plain) < ev_ord[h])
                is_new = false;
        }
        ev_tmp[i] = is_new ? 1u : 0u;
    }
    __syncthreads();
    uint32_t my_new = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        if (!ev_tmp[i])
            continue;   // already a member: add is a no-op, membership comes from the table
        my_new++;
        const unsigned long long ord = ev_ord[ev_used[i]];
        uint32_t rank = 0;
        for (uint32_t k = 0; k < n; k++)
            rank += (ev_tmp[k] && ev_ord[ev_used[k]] < ord) ? 1u : 0u;
        // ordinals of distinct keys can tie only if (buffer, j, phase) coincide, which
        // cannot happen: one decode adds at most one key.
        if (have + rank < (uint32_t)B200ADSB_ICAO_FILTER_SIZE)
            new_keys[rank] = ev_keys[ev_used[i]];
        else
            ev_tmp[i] = 2u;   // dropped: "icao24 hash table full"
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        if (ev_tmp[i] == 2u)
            ev_ord[ev_used[i]] = kNever;
    atomicAdd(&counters[C_NEWCNT], my_new);
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t nn = counters[C_NEWCNT];
        const uint32_t room = (uint32_t)B200ADSB_ICAO_FILTER_SIZE - have;
        counters[C_ADMIT] = min(nn, room);
        if (nn > room)
            atomicOr(&counters[C_FLAGS], F_FILTER_FULL);
    }
}

__global__ void __launch_bounds__(1024) events_finalize_kernel(const FinalizeArgs a)
{
    events_finalize_body(a.ev_keys, a.ev_ord, a.ev_used, a.ev_tmp, a.new_keys, a.counters, a.members, a.ev_mask,
                         a.bloom);
}

// after resolve: the admitted keys join the filter; the event table is recycled.
// A bad batch (candidate pool / event table / exchange buffer overflow here or on another rank) is
// not committed: nothing of it may reach the filter, the host redoes it.  Enqueue-only batches
// (d_result != nullptr) additionally honour and raise the sticky flag: once a queued batch failed,
// the batches queued behind it are not committed either -- they would otherwise run against a
// filter that lacks the failed batch's addresses and then leave their own in it, which no re-run
// could undo -- until the host acknowledges (b200adsb_async_acknowledge or any synchronous call).
// flip_tail: carry mode, the tail saved by this batch becomes the stream tail.
__device__ __forceinline__ void events_commit_body(uint32_t *ev_keys, unsigned long long *ev_ord,
                                                   const uint32_t *ev_used, const uint32_t *new_keys,
                                                   uint32_t *counters, uint32_t *members, uint32_t *d_result,
                                                   uint32_t cap, int flip_tail)
{
    const uint32_t fl = counters[C_FLAGS];
    const bool sticky = d_result != nullptr && counters[C_STICKY] != 0u;
    const bool bad = (fl & kBadMask) != 0u || sticky;
    const uint32_t n = counters[C_EV_USED], adm = bad ? 0u : counters[C_ADMIT];
    for (uint32_t i = threadIdx.x; i < adm; i += blockDim.x)
        members_insert(members, new_keys[i]);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t h = ev_used[i];
        ev_keys[h] = 0u;
        ev_ord[h] = kNever;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        counters[C_MEMBERS] += adm;
        counters[C_EV_USED] = 0;
        counters[C_ADMIT] = 0;
        counters[C_NEWCNT] = 0;
        if (flip_tail && !bad)
            counters[C_TAILCUR] ^= 1u;
        if (d_result) {       // enqueue-only batch outcome; the per-batch counters are cleared here
            d_result[0] = bad ? 0u : counters[C_FRAMES];
            d_result[1] = (fl & kBadMask) | (sticky ? F_SKIPPED : 0u);
            d_result[2] = counters[C_CAND];
            d_result[3] = counters[C_FRAMES] > cap ? 1u : 0u;
            if (bad)
                counters[C_STICKY] = 1u;
            counters[C_POOL] = 0;
            counters[C_FLAGS] = 0;
            counters[C_CAND] = 0;
        }
    }
}
struct CommitArgs {
    uint32_t *ev_keys;
    unsigned long long *ev_ord;
    const uint32_t *ev_used, *new_keys;
    uint32_t *counters, *members, *d_result;
    uint32_t cap;
    int flip_tail;
};
__global__ void __launch_bounds__(1024) events_commit_kernel(const CommitArgs a)
{
    events_commit_body(a.ev_keys, a.ev_ord, a.ev_used, a.new_keys, a.counters, a.members, a.d_result, a.cap,
                       a.flip_tail);
}

// ================================================================== resolve
struct ResolveParams {
    const uint32_t *rec;
    const uint2 *tile_dir;
    uint32_t n_tiles;
    int tiles_per_buffer;
    uint32_t *emit_info;       // per tile (at its pool base): compact list of emitting records
    uint32_t *tile_emit;       // per tile: frames emitted
    uint32_t *cta_sum;         // per resolve block (32 tiles): frames emitted
    const uint32_t *bloom;
    int32_t *rec_score;        // nullable: per record best score (diagnostics / message API)
    const uint32_t *members;
    const uint32_t *ev_keys;
    const unsigned long long *ev_ord;
    uint32_t ev_mask;
    unsigned long long ord_first, ord_stride;
};

// one warp per tile, 32 tiles per block; lanes stride over the tile's records
constexpr int kResolveThreads = 1024;
__device__ __forceinline__ void resolve_body(const ResolveParams &p, const uint32_t blk)
{
    __shared__ uint32_t s_cnt[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tile = blk * 32 + warp;
    uint32_t emitted = 0;
    if (tile < p.n_tiles) {
        const uint2 d = p.tile_dir[tile];
        const uint32_t b = tile / (uint32_t)p.tiles_per_buffer;
        const unsigned long long ord_buf = (p.ord_first + (unsigned long long)b * p.ord_stride) << 20;
        // one 24-byte record = three aligned 8-byte loads, fetched one round ahead (the rounds of a
        // tile are otherwise a chain of dependent DRAM latencies)
        uint2 na = make_uint2(0u, 0u), nb = na, nc = na;
        if ((uint32_t)lane < d.y) {
            const uint2 *r2 = reinterpret_cast<const uint2 *>(p.rec + 6ull * (d.x + (uint32_t)lane));
            na = r2[0]; nb = r2[1]; nc = r2[2];
        }
        for (uint32_t i = lane; i < d.y; i += 32) {
            // the five words are scored without a branch per word (lanes hold different kinds: a
            // switch here ran at 14 of 32 lanes)
            const uint2 ra = na, rb = nb, rc = nc;
            if (i + 32 < d.y) {
                const uint2 *r2 = reinterpret_cast<const uint2 *>(p.rec + 6ull * (d.x + i + 32));
                na = r2[0]; nb = r2[1]; nc = r2[2];
            }
            const uint32_t j = ra.x;
            const uint32_t w5[5] = {ra.y, rb.x, rb.y, rc.x, rc.y};
            int best = -2;                        // demod_2400.rs:152
            uint32_t best_t = 0, best_len = 7;
#pragma unroll
            for (int tt = 0; tt < 5; tt++) {
                const uint32_t wd = w5[tt];
                const uint32_t kind = wd >> 29, key = wd & 0xffffffu;
                // membership: address 0 always tests true (icao_filter.rs:71,78); the bloom word
                // rejects almost every other key with one load
                bool m = key == 0u;
                if (kind != K_NONE && key != 0u && bloom_hit(p.bloom, key))
                    m = members_has(p.members, key) ||
                        event_first(p.ev_keys, p.ev_ord, p.ev_mask, key) < (ord_buf | ((unsigned long long)j << 3) | (unsigned)tt);
                // mode_s/mod.rs:56-134 as selects: (member, not member) scores and the frame length
                const bool df1718 = kind == K_DF17 || kind == K_DF18;
                const int s_m = kind == K_DF11_IID0 ? 1600 : (df1718 ? 1800 : 1000);
                const int s_n = kind == K_DF11_IID0 ? 750 : (df1718 ? 1400 : (kind == K_PAR_LONG ? -2 : -1));
                int score = m ? s_m : s_n;
                if (kind == K_NONE)
                    score = -3;                   // None / rejected statelessly: never beats -2
                const uint32_t len = (df1718 || kind == K_PAR_LONG) ? 14u : 7u;
                if (score > best) {               // demod_2400.rs:185 (strict)
                    best = score;
                    best_t = (uint32_t)tt;
                    best_len = len;
                }
            }
            const bool emit = best >= 0;          // demod_2400.rs:203
            if (p.rec_score)
                p.rec_score[d.x + i] = best;
            // the tile's emitting records, compacted in ascending j at the front of its range:
            // score<<17 | record index in tile<<4 | long<<3 | try_phase-4
            const unsigned act = __activemask();
            const unsigned em = __ballot_sync(act, emit);
            if (emit)
                p.emit_info[d.x + emitted + (uint32_t)__popc(em & ((1u << lane) - 1u))] =
                    ((uint32_t)best << 17) | (i << 4) | ((best_len == 14u ? 1u : 0u) << 3) | best_t;
            emitted += (uint32_t)__popc(em);
        }
        if (lane == 0)
            p.tile_emit[tile] = emitted;
    }
    if (lane == 0)
        s_cnt[warp] = emitted;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = s_cnt[lane];
#pragma unroll
        for (int o = 16; o; o >>= 1)
            v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0)
            p.cta_sum[blk] = v;
    }
    __syncthreads();   // s_cnt is reused by the next block of tiles in the fused small-batch kernel
}
__device__ __forceinline__ void tile_scan_body(uint32_t *tile_emit, uint32_t n_tiles, uint32_t *counters);
// the last block to finish scans the per-block sums (saves a dependent launch)
__global__ void __launch_bounds__(kResolveThreads) resolve_kernel(const ResolveParams p, uint32_t *counters)
{
    __shared__ uint32_t s_last;
    resolve_body(p, blockIdx.x);
    if (threadIdx.x == 0) {
        __threadfence();                                  // cta_sum[blk] before the ticket
        s_last = atomicAdd(&counters[C_TICKET], 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        tile_scan_body(p.cta_sum, gridDim.x, counters);   // -> exclusive prefix, counters[C_FRAMES]
        if (threadIdx.x == 0)
            counters[C_TICKET] = 0;
    }
}

// exclusive scan of tile_emit (in place) by one block; total -> counters[C_FRAMES];
// per-buffer counts optional
__device__ __forceinline__ void tile_scan_body(uint32_t *tile_emit, uint32_t n_tiles, uint32_t *counters)
{
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0)
        s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_tiles ? *reinterpret_cast<volatile uint32_t *>(tile_emit + i) : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        if (lane == 31)
            s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t x = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o)
                    x += t;
            }
            s_w[lane] = x;
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t excl = carry + (warp ? s_w[warp - 1] : 0u) + incl - v;
        if (i < n_tiles)
            tile_emit[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023)
            s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        counters[C_FRAMES] = s_carry;
}
__global__ void __launch_bounds__(1024) tile_scan_kernel(uint32_t *tile_emit, uint32_t n_tiles, uint32_t *counters)
{
    tile_scan_body(tile_emit, n_tiles, counters);
}

__global__ void buffer_counts_kernel(const uint32_t *tile_cnt, uint32_t n_buffers, int tpb, uint32_t *out)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buffers)
        return;
    uint32_t v = 0;
    for (int k = 0; k < tpb; k++)
        v += tile_cnt[(size_t)b * tpb + k];
    out[b] = v;
}

// ================================================================== emit
struct EmitParams {
    const void *in;
    const uint32_t *lengths;
    uint32_t spb;
    unsigned long long stride;
    const uint32_t *rec;
    const uint2 *tile_dir;
    const uint32_t *emit_info;
    const uint32_t *tile_cnt;    // frames per tile
    const uint32_t *cta_excl;    // exclusive prefix per block of 32 tiles
    int carry;
    const uint32_t *tails;     // carry mode: the two stream tails
    const uint32_t *counters;  //             and [C_TAILCUR]
    uint32_t n_tiles;
    int tiles_per_buffer;
    b200adsb_frame *out;
    uint32_t cap;
    const uint8_t *msgs;   // message-level API: frames copy their bytes from here
};

// one warp per tile; every emitted frame is re-sliced from the input by the whole warp
// (src/demod_2400.rs:158-182 in closed form: bit n of try_phase t is decided at
//  P = 5(j+19)+t+12n, sample P/5, correlator P%5)
template <bool FROM_MAG>
__device__ __forceinline__ void emit_body(const EmitParams &p, const uint32_t blk)
{
    __shared__ uint16_t s_mag[32][288];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tile = blk * 32 + warp;
    // this tile's first output slot: block prefix + the counts of the block's earlier tiles
    const uint32_t t_l = blk * 32 + (uint32_t)lane;
    const uint32_t c_l = t_l < p.n_tiles ? p.tile_cnt[t_l] : 0u;
    uint32_t incl = c_l;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += t;
    }
    const uint32_t my_cnt = __shfl_sync(0xffffffffu, c_l, warp);
    uint32_t out_idx = p.cta_excl[blk] + __shfl_sync(0xffffffffu, incl - c_l, warp);
    if (tile >= p.n_tiles || my_cnt == 0u)
        return;
    const uint2 d = p.tile_dir[tile];
    const uint32_t b = tile / (uint32_t)p.tiles_per_buffer;
    const int len = p.lengths ? (int)min(p.lengths[b], p.spb) : (int)p.spb;
    const CarrySrc cs{reinterpret_cast<const uint32_t *>(p.in), p.stride, p.lengths, p.spb, p.tails, p.counters,
                      (FROM_MAG || p.msgs) ? 0 : p.carry};
    for (uint32_t base = 0; base < my_cnt; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t info = i < my_cnt ? p.emit_info[d.x + i] : 0u;
        const uint32_t jmine = i < my_cnt ? p.rec[6ull * (d.x + ((info >> 4) & 0x1fffu))] : 0u;
        const int nthis = (int)min(32u, my_cnt - base);
        for (int src = 0; src < nthis; src++) {
            const uint32_t inf = __shfl_sync(0xffffffffu, info, src);
            const uint32_t j = __shfl_sync(0xffffffffu, jmine, src);
            const int t = 4 + (int)(inf & 7u), flen = (inf & 8u) ? 14 : 7;
            uint32_t words[4] = {0, 0, 0, 0};
            if (p.msgs == nullptr) {
                // magnitudes of data[j+19 .. j+19+288)
                for (int k = lane; k < 288; k += 32) {
                    const int idx = (int)j + 19 + k;
                    uint32_t m = 0;
                    if (FROM_MAG) {
                        const uint16_t *dd = reinterpret_cast<const uint16_t *>(p.in) + (unsigned long long)b * p.stride;
                        if (idx < kMagLen)
                            m = dd[idx];
                    } else {
                        const uint32_t *bb = reinterpret_cast<const uint32_t *>(p.in) + (unsigned long long)b * p.stride;
                        const int s = idx - kTrailing;
                        // == mag_pair (exhaustively checked)
                        m = mag_bits_fast(iq_word(bb, s, len, cs, b)) & 0xffffu;
                    }
                    s_mag[warp][k] = (uint16_t)m;
                }
                __syncwarp();
#pragma unroll
                for (int wi = 0; wi < 4; wi++) {
                    const int n = 32 * wi + lane;
                    bool one = false;
                    if (n < 112) {
                        const int P = t + 12 * n;          // relative to 5*(j+19)
                        const int i5 = P / 5, phi = P - 5 * i5;
                        const uint16_t *m = &s_mag[warp][i5];
                        const int m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3];
                        int x;
                        switch (phi) {                      // demod_2400.rs:72-83
                        case 0: x = 5 * m0 - 3 * m1 - 2 * m2; break;
                        case 1: x = 4 * m0 - m1 - 3 * m2; break;
                        case 2: x = 3 * m0 + m1 - 4 * m2; break;
                        case 3: x = 2 * m0 + 3 * m1 - 5 * m2; break;
                        default: x = m0 + 5 * m1 - 5 * m2 - m3; break;
                        }
                        one = x > 0;
                    }
                    words[wi] = __ballot_sync(0xffffffffu, one);
                }
                __syncwarp();
            }
            if (out_idx < p.cap) {
                // frame = 7 u32 words: msg[0..13], len, phase | score, reserved | j | buffer
                uint8_t by[16];
#pragma unroll
                for (int kb = 0; kb < 14; kb++) {
                    uint32_t v;
                    if (p.msgs)
                        v = p.msgs[14ull * ((unsigned long long)b * 1024ull + j) + kb];
                    else
                        v = __brev((words[kb >> 2] >> (8 * (kb & 3))) & 0xffu) >> 24;
                    by[kb] = (kb < flen) ? (uint8_t)v : (uint8_t)0;
                }
                by[14] = (uint8_t)flen;
                by[15] = (uint8_t)t;
                if (lane == 0) {
                    uint32_t *o = reinterpret_cast<uint32_t *>(p.out + out_idx);
#pragma unroll
                    for (int wq = 0; wq < 4; wq++)
                        o[wq] = by[4 * wq] | (by[4 * wq + 1] << 8) | (by[4 * wq + 2] << 16) |
                                ((uint32_t)by[4 * wq + 3] << 24);
                    o[4] = (inf >> 17) & 0x7fffu;   // score (>= 0 here), reserved = 0
                    o[5] = j;
                    o[6] = b;
                }
            }
            out_idx++;
        }
    }
}

// Large batches: one warp per FRAME (grid-stride over the batch's frames), so that tiles holding many
// frames (dense traffic) do not serialise in one warp.  Frame f belongs to the tile found by a
// binary search over the per-block prefix (32 tiles per resolve block) and a warp scan of that
// block's 32 tile counts; output slot = f, i.e. (buffer, j) order as before.
constexpr int kEmitWarps = 8;
// The commit step has no data dependency on the emit step (it touches the filter, the event table and
// other counter words), so it rides along as the last block of this launch instead of a launch of
// its own.
__device__ __forceinline__ void commit_tail(const CommitArgs &a)
{
    events_commit_body(a.ev_keys, a.ev_ord, a.ev_used, a.new_keys, a.counters, a.members, a.d_result, a.cap,
                       a.flip_tail);
}
template <bool FROM_MAG>
__global__ void __launch_bounds__(32 * kEmitWarps) emit_frames_kernel(const EmitParams p, const uint32_t *counters,
                                                                      const uint32_t n_ctas, const CommitArgs ca)
{
    __shared__ uint16_t s_mag[kEmitWarps][288];
    if (blockIdx.x == gridDim.x - 1) {          // the extra block
        if (ca.counters)
            commit_tail(ca);
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t total = min(counters[C_FRAMES], p.cap);
    const uint32_t nwarps = (gridDim.x - 1) * kEmitWarps;
    for (uint32_t f = blockIdx.x * kEmitWarps + warp; f < total; f += nwarps) {
        // resolve block whose exclusive prefix is the last one <= f: 32-ary search, one probe per lane
        // (two dependent loads for up to 1024 blocks instead of ten)
        uint32_t lo = 0, span = n_ctas;         // the answer lies in [lo, lo + span)
        while (span > 1) {
            const uint32_t stp = (span + 31) / 32;
            const uint32_t idx = lo + (uint32_t)lane * stp;
            const bool le = idx < lo + span && p.cta_excl[idx] <= f;
            const int seg = __popc(__ballot_sync(0xffffffffu, le)) - 1;   // lane 0 always holds (cta_excl[lo] <= f)
            const uint32_t nlo = lo + (uint32_t)seg * stp;
            span = min(stp, lo + span - nlo);
            lo = nlo;
        }
        const uint32_t blk = lo, rel = f - p.cta_excl[blk];
        const uint32_t t_l = blk * 32 + (uint32_t)lane;
        const uint32_t c_l = t_l < p.n_tiles ? p.tile_cnt[t_l] : 0u;
        uint32_t incl = c_l;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        const unsigned owner = __ballot_sync(0xffffffffu, rel >= incl - c_l && rel < incl);
        const int wl = __ffs(owner) - 1;        // exactly one lane owns rel (rel < block total)
        const uint32_t tile = blk * 32 + (uint32_t)wl;
        const uint32_t i = rel - __shfl_sync(0xffffffffu, incl - c_l, wl);
        const uint2 d = p.tile_dir[tile];
        const uint32_t b = tile / (uint32_t)p.tiles_per_buffer;
        const int len = p.lengths ? (int)min(p.lengths[b], p.spb) : (int)p.spb;
        const CarrySrc cs{reinterpret_cast<const uint32_t *>(p.in), p.stride, p.lengths, p.spb, p.tails, p.counters,
                          (FROM_MAG || p.msgs) ? 0 : p.carry};
        const uint32_t inf = p.emit_info[d.x + i];
        const uint32_t j = p.rec[6ull * (d.x + ((inf >> 4) & 0x1fffu))];
        const int t = 4 + (int)(inf & 7u), flen = (inf & 8u) ? 14 : 7;
        uint32_t words[4] = {0, 0, 0, 0};
        if (p.msgs == nullptr) {
            for (int k = lane; k < 288; k += 32) {      // magnitudes of data[j+19 .. j+19+288)
                const int idx = (int)j + 19 + k;
                uint32_t m = 0;
                if (FROM_MAG) {
                    const uint16_t *dd = reinterpret_cast<const uint16_t *>(p.in) + (unsigned long long)b * p.stride;
                    if (idx < kMagLen)
                        m = dd[idx];
                } else {
                    const uint32_t *bb = reinterpret_cast<const uint32_t *>(p.in) + (unsigned long long)b * p.stride;
                    m = mag_bits_fast(iq_word(bb, idx - kTrailing, len, cs, b)) & 0xffffu;   // == mag_pair
                }
                s_mag[warp][k] = (uint16_t)m;
            }
            __syncwarp();
#pragma unroll
            for (int wi = 0; wi < 4; wi++) {
                const int n = 32 * wi + lane;
                bool one = false;
                if (n < 112) {
                    const int P = t + 12 * n;          // relative to 5*(j+19)
                    const int i5 = P / 5, phi = P - 5 * i5;
                    const uint16_t *m = &s_mag[warp][i5];
                    const int m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3];
                    int x;
                    switch (phi) {                      // demod_2400.rs:72-83
                    case 0: x = 5 * m0 - 3 * m1 - 2 * m2; break;
                    case 1: x = 4 * m0 - m1 - 3 * m2; break;
                    case 2: x = 3 * m0 + m1 - 4 * m2; break;
                    case 3: x = 2 * m0 + 3 * m1 - 5 * m2; break;
                    default: x = m0 + 5 * m1 - 5 * m2 - m3; break;
                    }
                    one = x > 0;
                }
                words[wi] = __ballot_sync(0xffffffffu, one);
            }
            __syncwarp();
        }
        uint8_t by[16];
#pragma unroll
        for (int kb = 0; kb < 14; kb++) {
            uint32_t v;
            if (p.msgs)
                v = p.msgs[14ull * ((unsigned long long)b * 1024ull + j) + kb];
            else
                v = __brev((words[kb >> 2] >> (8 * (kb & 3))) & 0xffu) >> 24;
            by[kb] = (kb < flen) ? (uint8_t)v : (uint8_t)0;
        }
        by[14] = (uint8_t)flen;
        by[15] = (uint8_t)t;
        if (lane == 0) {
            uint32_t *o = reinterpret_cast<uint32_t *>(p.out + f);
#pragma unroll
            for (int wq = 0; wq < 4; wq++)
                o[wq] = by[4 * wq] | (by[4 * wq + 1] << 8) | (by[4 * wq + 2] << 16) | ((uint32_t)by[4 * wq + 3] << 24);
            o[4] = (inf >> 17) & 0x7fffu;   // score (>= 0 here), reserved = 0
            o[5] = j;
            o[6] = b;
        }
    }
}

// Small batches (one SDR read per call): the whole second stage in one launch of one block --
// finalise, resolve, scan, emit, commit -- because five dependent launches of a few
// microseconds each would dominate the call.
template <bool FROM_MAG>
__global__ void __launch_bounds__(kResolveThreads) resolve_small_kernel(const FinalizeArgs fa, const ResolveParams rp,
                                                                        const EmitParams ep, const uint32_t n_ctas,
                                                                        const int flip_tail)
{
    events_finalize_body(fa.ev_keys, fa.ev_ord, fa.ev_used, fa.ev_tmp, fa.new_keys, fa.counters, fa.members,
                         fa.ev_mask, fa.bloom);
    __syncthreads();
    for (uint32_t blk = 0; blk < n_ctas; blk++)
        resolve_body(rp, blk);
    __syncthreads();
    tile_scan_body(rp.cta_sum, n_ctas, fa.counters);
    __syncthreads();
    for (uint32_t blk = 0; blk < n_ctas; blk++) {
        emit_body<FROM_MAG>(ep, blk);
        __syncthreads();
    }
    events_commit_body(fa.ev_keys, fa.ev_ord, fa.ev_used, fa.new_keys, fa.counters, fa.members, nullptr, 0u, flip_tail);
}

// carry mode: the last 326 samples of the stream (walking back over this batch's buffers,
// then into the old tail) become the next batch's leading samples
__global__ void save_tail_kernel(const uint32_t *in, unsigned long long stride, const uint32_t *lengths,
                                 uint32_t spb, uint32_t n_buffers, uint32_t *tails, const uint32_t *counters)
{
    const int k = threadIdx.x;
    if (k >= kTrailing)
        return;
    const uint32_t cur = counters[C_TAILCUR] & 1u;
    const uint32_t *old_tail = tails + kTailWords * cur;
    uint32_t *new_tail = tails + kTailWords * (cur ^ 1u);   // becomes current when the batch commits
    int back = kTrailing - 1 - k;          // 0 = the very last sample of the stream
    uint32_t w = 0;
    bool found = false;
    for (int b = (int)n_buffers - 1; b >= 0 && !found; b--) {
        const int len = lengths ? (int)min(lengths[b], spb) : (int)spb;
        if (back < len) {
            w = in[(unsigned long long)b * stride + (unsigned)(len - 1 - back)];
            found = true;
        } else {
            back -= len;
        }
    }
    if (!found && back < kTrailing)
        w = old_tail[kTrailing - 1 - back];
    new_tail[k] = w;
}

// ================================================================== frame gather (sharded streams)
// The reference emits ONE stream of frames in (buffer, j) order (dump1090_rs/src/main.rs:166-200).
// A sharded run leaves every rank with the frames of its own buffers (local buffer indices, ascending
// (buffer, j)); they travel as fixed-size blocks -- row 0 = {count, 0, ...}, rows 1.. = frames -- through
// one all-gather, and every rank (or only the emitting one) merges them:
// frames_pack_kernel: this rank's block.  count: *d_count if given, else `count`.
__global__ void frames_pack_kernel(const b200adsb_frame *frames, const uint32_t *d_count, uint32_t count,
                                   b200adsb_frame *block, uint32_t rows_cap)
{
    const uint32_t n = d_count ? *d_count : count;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        uint32_t *h = reinterpret_cast<uint32_t *>(block);
        h[0] = n;
        for (int k = 1; k < 7; k++)
            h[k] = 0;
    }
    const uint32_t m = min(n, rows_cap);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(frames);
    uint32_t *dst = reinterpret_cast<uint32_t *>(block + 1);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < 7u * m; i += gridDim.x * blockDim.x)
        dst[i] = src[i];
}
// The same gather fused with its transport, like the event exchange above: every rank owns
//   u64 flags[2][world]   then   frame blocks[2][world][1 + rows_cap]
// in symmetric memory; frames_push_symm_kernel stores this rank's block into slot [parity][rank] of every
// rank's buffer over NVLink and the last block to finish raises the rank's flag everywhere;
// frames_merge_kernel, given the local flags and the epoch, waits for all of them before it merges.
__device__ __forceinline__ b200adsb_frame *symm_frames(unsigned char *base, uint32_t world, uint32_t rows_cap,
                                                       uint32_t parity, uint32_t r)
{
    return reinterpret_cast<b200adsb_frame *>(base + 16ull * world) +
           ((unsigned long long)parity * world + r) * ((unsigned long long)rows_cap + 1);
}
__global__ void frames_push_symm_kernel(const b200adsb_frame *frames, const uint32_t *d_count, uint32_t count,
                                        unsigned char *const *peer_bufs, uint32_t rank, uint32_t world,
                                        uint32_t rows_cap, unsigned long long epoch, uint32_t *ticket)
{
    __shared__ uint32_t s_last;
    const uint32_t parity = (uint32_t)(epoch & 1ull);
    const uint32_t n = d_count ? *d_count : count;
    const uint32_t m = min(n, rows_cap);
    if (blockIdx.x == 0 && threadIdx.x < world) {
        uint32_t *h = reinterpret_cast<uint32_t *>(symm_frames(peer_bufs[threadIdx.x], world, rows_cap, parity, rank));
        h[0] = n;
        for (int k = 1; k < 7; k++)
            h[k] = 0;
    }
    const uint32_t *src = reinterpret_cast<const uint32_t *>(frames);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < 7u * m; i += gridDim.x * blockDim.x) {
        const uint32_t v = src[i];
        for (uint32_t q = 0; q < world; q++)
            reinterpret_cast<uint32_t *>(symm_frames(peer_bufs[q], world, rows_cap, parity, rank) + 1)[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if (threadIdx.x < world) {
            unsigned long long *flag =
                reinterpret_cast<unsigned long long *>(peer_bufs[threadIdx.x]) + (unsigned long long)parity * world + rank;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
        }
        if (threadIdx.x == 0)
            *ticket = 0;
    }
}

__device__ __forceinline__ unsigned long long frame_key(const b200adsb_frame *f, uint32_t rank, uint32_t world)
{
    return ((unsigned long long)(f->buffer * world + rank) << 32) | f->j;   // round-robin dealing: g = local * world + rank
}
// frames_merge_kernel: n_ranks blocks of (1 + rows_cap) rows -> out[] in global (buffer, j) order with global
// buffer indices.  A frame's slot = its index in its own list + the number of frames with a smaller key in
// every other list (binary search; keys of different ranks never tie: they are different buffers).
// n_out[0] = total frames, n_out[1] = 1 if a rank had more frames than rows_cap or the total exceeds cap.
// wait_flags != nullptr: the blocks arrive by peer stores; wait (bounded, ~30 s) until the n_ranks flags have
// reached `epoch`.  A peer that never arrives sets n_out[1] = 3.
__global__ void frames_merge_kernel(const b200adsb_frame *gathered, uint32_t n_ranks, uint32_t rows_cap,
                                    b200adsb_frame *out, uint32_t cap, uint32_t *n_out,
                                    const unsigned long long *wait_flags, unsigned long long epoch)
{
    if (wait_flags) {
        __shared__ uint32_t s_timeout;
        if (threadIdx.x == 0)
            s_timeout = 0;
        __syncthreads();
        if (threadIdx.x < n_ranks) {
            const long long t0 = clock64();
            for (;;) {
                unsigned long long v;
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(wait_flags + threadIdx.x) : "memory");
                if (v >= epoch)
                    break;
                if (clock64() - t0 > 60000000000ll) {
                    s_timeout = 1;
                    break;
                }
                __nanosleep(200);
            }
        }
        __syncthreads();
        if (s_timeout) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                n_out[0] = 0;
                n_out[1] = 3;
            }
            return;
        }
    }
    const size_t stride = (size_t)rows_cap + 1;
    uint32_t total = 0, ovf = 0;
    for (uint32_t r = 0; r < n_ranks; r++) {
        const uint32_t c = *reinterpret_cast<const uint32_t *>(gathered + r * stride);
        ovf |= c > rows_cap ? 1u : 0u;
        total += min(c, rows_cap);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        n_out[0] = total;
        n_out[1] = (ovf || total > cap) ? 1u : 0u;
    }
    uint32_t base = 0;
    for (uint32_t r = 0; r < n_ranks; r++) {
        const b200adsb_frame *mine = gathered + r * stride + 1;
        const uint32_t cnt = min(*reinterpret_cast<const uint32_t *>(gathered + r * stride), rows_cap);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
            const unsigned long long key = frame_key(mine + i, r, n_ranks);
            uint32_t pos = i;
            for (uint32_t q = 0; q < n_ranks; q++) {
                if (q == r)
                    continue;
                const b200adsb_frame *other = gathered + q * stride + 1;
                uint32_t lo = 0, hi = min(*reinterpret_cast<const uint32_t *>(gathered + q * stride), rows_cap);
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (frame_key(other + mid, q, n_ranks) < key)
                        lo = mid + 1;
                    else
                        hi = mid;
                }
                pos += lo;
            }
            if (pos < cap) {
                b200adsb_frame f = mine[i];
                f.buffer = f.buffer * n_ranks + r;
                out[pos] = f;
            }
        }
        base += cnt;
    }
    (void)base;
}

// ================================================================== filter helpers
__global__ void filter_add_kernel(uint32_t *members, uint32_t *counters, uint32_t key)
{
    if (threadIdx.x || blockIdx.x)
        return;
    if (key == 0u || members_has(members, key))
        return;
    if (counters[C_MEMBERS] >= (uint32_t)B200ADSB_ICAO_FILTER_SIZE) {
        atomicOr(&counters[C_FLAGS], F_FILTER_FULL);
        return;
    }
    members_insert(members, key);
    counters[C_MEMBERS] += 1;
}
__global__ void filter_test_kernel(const uint32_t *members, uint32_t key, uint32_t *out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *out = (key == 0u || members_has(members, key)) ? 1u : 0u;
}
__global__ void filter_restore_kernel(uint32_t *members, uint32_t *counters, const uint32_t *keys,
                                      uint32_t n)
{
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
        if (keys[i])
            members_insert(members, keys[i]);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t c = 0;
        for (uint32_t h = 0; h < kMemberSlots; h++)
            c += members[h] != 0u;
        counters[C_MEMBERS] = c;
    }
}
__global__ void fill_u64_kernel(unsigned long long *p, size_t n, unsigned long long v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = v;
}

}  // namespace b200
