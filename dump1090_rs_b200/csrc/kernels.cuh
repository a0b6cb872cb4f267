// dump1090_rs_b200/csrc/kernels.cuh -- device side of libb200adsb (sm_100a).
//
// The reference scans one sample at a time on one CPU thread
// (src/demod_2400.rs:115-212).  Here the same arithmetic is re-organised for a GPU.
// This file holds what both generations of the stage-1 kernel share (exact fast magnitude,
// CRC-24 by fields, classification, event table, gates), the previous generation itself
// (scan_kernel, "v6", selectable with B200ADSB_SCAN=6 for A/B runs; the current one is
// scan7_kernel in scan7.cuh), stage 2 and the API-parity kernels.
//
//   scan_kernel (one thread block per tile of T output positions)
//     P1  IQ -> u16 magnitude, edge bits and first differences into shared memory
//                                                          (src/utils.rs:43-58)
//     P2  every sample's five PPM correlator signs as bit planes de-interleaved modulo 12
//         samples (one Mode-S bit period at 2.4 Msps is 12/5 samples, so message bit n of
//         try-phase t lives at 1/5-sample position P = 5(j+19)+t+12n: sample P/5,
//         correlator P%5)                                  (src/demod_2400.rs:62-83,158-182)
//     P3  the five preamble templates evaluated 32 positions at a time as AND/shift
//         of the edge bitmaps, then SNR + quiet-zone gates on the matches
//                                                          (src/demod_2400.rs:127-146,215-321)
//     P4  per surviving position and try-phase: five 23-bit field extracts from
//         the planes, DF, CRC-24 syndrome by table (GF(2)-linear), stateless
//         classification -> one 24-byte record; ICAO add-events by atomicMin
//                                                          (src/mode_s/mod.rs:34-139, src/crc.rs:263-282)
//   finalize / resolve / emit / commit kernels
//         the sequential ICAO filter (src/icao_filter.rs) evaluated order-free:
//         member(a) at ordinal o  <=>  a == 0 || a in filter before the batch ||
//         firstAdd(a) < o; best-of-5 with the reference's strict '>' rule;
//         frames written in (buffer, j) order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200adsb.h"

// build-time experiment knobs (scripts/ab.sh compiles variants with -D...)
#ifndef B200_SCAN_MIN_BLOCKS
#define B200_SCAN_MIN_BLOCKS 5
#endif
#ifndef B200_AGG_ATOMICS
#define B200_AGG_ATOMICS 1
#endif

namespace b200 {

constexpr int kTrailing = B200ADSB_TRAILING_SAMPLES;          // lib.rs:24
constexpr int kMaxSamples = B200ADSB_MODES_MAG_BUF_SAMPLES;    // lib.rs:22
constexpr int kMagLen = B200ADSB_MAG_DATA_LEN;
constexpr int kHaloFront = 2;    // tile mag index 0 <-> data index tile_start-2 (16 B aligned IQ loads)
constexpr int kHaloTot = 296;    // mags needed per tile = T + 296 (max tap j+289, +2 front, padded to 8)
constexpr int kStep = 384;       // 12 residues x 32 lanes: samples per warp step
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunk = 248;      // new pair-slots per warp iteration of P1 (8 per lane, lane 31 overlaps)
constexpr int kHalf = 192;       // half block: 12 residues x 16 lanes-steps
constexpr int kQuarter = 96;     // P2 task = 8 samples at stride 12 of one residue class
constexpr int kQuarterPad = 108; // 96 pair-slots + 12 mirrored from the next quarter block (bank skew 12)
constexpr int kQueueCap = 256;   // template matches per template case awaiting the gates (overflow: in place)
constexpr int kCandCap = 176;    // survivors decoded per window
constexpr int kFieldItems = 5 * kCandCap;   // (survivor, try_phase) items whose fields are staged
constexpr int kLutWords = 12 * 25;          // field-extraction table: (residue of j+19, try_phase, field)
constexpr int kMaxTile = 8184;   // tile mag indices (< T+2) fit 13 bits; surv words <= 256
constexpr int kDefaultTile = 7384;   // 20 blocks of 384: 16 P1 chunks (2 per warp), 240 P2 tasks, 231 P3 words
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
       C_ADMIT = 6, C_NEWCNT = 7, C_TICKET = 8, C_WORDS = 16 };
enum : uint32_t { F_POOL_OVF = 1, F_EV_OVF = 2, F_FILTER_FULL = 4 };

constexpr uint32_t kMemberSlots = 8192;   // open addressing, >= 2 x 4096 keys
constexpr unsigned long long kNever = ~0ull;

struct ScanParams {
    const void *in;            // int16 (re,im) pairs, or u16 MagnitudeBuffer.data
    const uint32_t *lengths;   // nullable per-buffer sample counts
    uint32_t n_buffers;
    uint32_t spb;              // samples per buffer (length when lengths == NULL)
    unsigned long long stride; // IQ: samples between buffers; MAG: u16 elements
    int T, tiles_per_buffer;
    int vec_ok;                // 16-byte aligned base and stride % 4 == 0
    uint32_t *rec;             // pool of 6-word records {j, w[5]}
    uint32_t pool_cap;
    uint2 *tile_dir;           // per tile: (pool base, count)
    uint32_t *counters;
    uint32_t *ev_keys;
    unsigned long long *ev_ord;
    uint32_t *ev_used;
    uint32_t ev_mask;
    unsigned long long ord_first, ord_stride;
    const uint32_t *crc_tabs;  // CRC-24 field tables (global memory, L1 resident)
    const uint32_t *lut;       // [12][5][5] field extraction table for this tile size
    // shared memory layout of ScanSmem(T), computed once on the host
    uint32_t off_dd, off_planes, off_edges, edge_bytes, off_surv, off_queue, off_cand;
    int WP, nw;
    // opt-in stream continuity (B200ADSB_OPT_CARRY): the 326 leading MagnitudeBuffer slots of a
    // buffer hold the previous buffer's last samples instead of zeros
    int carry;
    const uint32_t *tail;      // last 326 IQ samples of the stream before this batch
};

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// shared memory plan of the scan kernel for tile size T.
// The halo-extended tile is H blocks of 384 samples = 2H half blocks of 192; the first H half
// blocks form stream A, the last H stream B, and everything downstream of the magnitude works
// on (A, B) sample pairs in the packed f32x2 pipe.
struct ScanSmem {
    int steps, MP, MPc, WP, nw, dd_words, edge_bytes;
    size_t off_dd, off_planes, off_edges, off_surv, off_tabs, off_queue, off_cand, off_lut, bytes;
    __host__ __device__ explicit ScanSmem(int T)
    {
        steps = (T + kHaloTot + kStep - 1) / kStep;   // H: 384-sample blocks
        MP = steps * kStep;
        MPc = MP + 64;                                // u16 magnitudes incl. the extra pair-slots
        WP = steps + 1;
        nw = (T + 31) / 32;
        size_t o = (size_t)(MPc + 8) * 2;
        o = (o + 15) & ~(size_t)15;
        off_dd = o;                                   // float2 first differences, 204 pair-slots per half block
        dd_words = 2 * kQuarterPad * (2 * ((steps + 1) / 2) + 1);   // half of the tile per P1/P2 pass
        if (dd_words < kFieldItems * 5)               // later reused as the P4 field buffer
            dd_words = kFieldItems * 5;
        o += (size_t)dd_words * 4;
        off_planes = o;                               // S[phi][rho][word], de-interleaved mod 12
        o += (size_t)5 * 12 * WP * 4;
        edge_bytes = round_up(MPc / 8 + 16, 16);
        o = (o + 15) & ~(size_t)15;
        off_edges = o;                                // R then F: one bit per sample, consecutive
        o += (size_t)2 * edge_bytes;
        off_surv = o;
        o += (size_t)nw * 4;
        off_tabs = off_lut = 0;                       // tables live in global memory (L1)
        off_queue = o;
        o += (size_t)5 * kQueueCap * 2;
        off_cand = o;
        o += (size_t)kCandCap * 2;
        bytes = (o + 15) & ~(size_t)15;
    }
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
__device__ __forceinline__ uint32_t a112(const uint32_t *t, uint32_t f)
{
    return __ldg(t + (f & 0xff)) ^ __ldg(t + 256 + ((f >> 8) & 0xff)) ^ __ldg(t + 512 + ((f >> 16) & 0x3f));
}
__device__ __forceinline__ uint32_t a56(const uint32_t *t, uint32_t f)
{
    return __ldg(t + kTab56 + (f & 0xff)) ^ __ldg(t + kTab56 + 256 + ((f >> 8) & 7));
}
__device__ __forceinline__ uint32_t syn112_fields(const uint32_t *t, const uint32_t f[5])
{
    uint32_t s = a112(t, f[0]);
    s = mulx(s) ^ a112(t, f[1]);
    s = mulx(s) ^ a112(t, f[2]);
    s = mulx(s) ^ a112(t, f[3]);
    s = mulx(s) ^ a112(t, f[4]);
    return s ^ (((f[0] >> 22) & 1u) << 1) ^ ((f[1] >> 22) & 1u);   // bits 110, 111
}
__device__ __forceinline__ uint32_t syn56_fields(const uint32_t *t, const uint32_t f[5])
{
    uint32_t s = a56(t, f[0]);
    s = mulx(s) ^ a56(t, f[1]);
    s = mulx(s) ^ a56(t, f[2]);
    s = mulx(s) ^ a56(t, f[3]);
    s = mulx(s) ^ a56(t, f[4]);
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

// eight consecutive IQ words of a buffer starting at sample s (zero outside [0, len):
// magnitude(0, 0) = 0 is exactly the MagnitudeBuffer zero fill, lib.rs:36-44)
__device__ __forceinline__ uint32_t iq_word(const uint32_t *b32, int s, int len, const uint32_t *prev, int prev_len)
{
    if (s >= 0)
        return s < len ? __ldg(b32 + s) : 0u;
    const int k = prev_len + s;          // carry mode: sample s < 0 lives in the previous buffer
    return (prev != nullptr && k >= 0) ? __ldg(prev + k) : 0u;
}
__device__ __forceinline__ void load_iq8(const uint32_t *b32, int s, int len, int vec_ok, const uint32_t *prev,
                                         int prev_len, uint32_t w[8])
{
    if (s >= 0 && s + 7 < len && vec_ok) {
        const int4 v0 = __ldg(reinterpret_cast<const int4 *>(b32 + s));
        const int4 v1 = __ldg(reinterpret_cast<const int4 *>(b32 + s + 4));
        w[0] = (uint32_t)v0.x; w[1] = (uint32_t)v0.y; w[2] = (uint32_t)v0.z; w[3] = (uint32_t)v0.w;
        w[4] = (uint32_t)v1.x; w[5] = (uint32_t)v1.y; w[6] = (uint32_t)v1.z; w[7] = (uint32_t)v1.w;
    } else {
#pragma unroll
        for (int e = 0; e < 8; e++)
            w[e] = iq_word(b32, s + e, len, prev, prev_len);
    }
}
// the buffer whose tail supplies the samples before buffer b in carry mode
__device__ __forceinline__ const uint32_t *carry_source(const void *in, unsigned long long stride,
                                                        const uint32_t *lengths, uint32_t spb, uint32_t b,
                                                        int carry, const uint32_t *tail, int *prev_len)
{
    *prev_len = 0;
    if (!carry)
        return nullptr;
    if (b == 0) {
        *prev_len = kTrailing;
        return tail;
    }
    *prev_len = lengths ? (int)min(lengths[b - 1], spb) : (int)spb;
    return reinterpret_cast<const uint32_t *>(in) + (unsigned long long)(b - 1) * stride;
}

// SNR and quiet-zone gates of one template match (demod_2400.rs:129,135-146) with the
// template's high/signal/noise (demod_2400.rs:226-317); cs = template case 0..4.
__device__ __forceinline__ void gate_eval(const uint16_t *mag, uint32_t *surv, int mi, uint32_t cs)
{
    const uint16_t *pp = mag + mi;
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
    const int jl = mi - kHaloFront;
    atomicOr(&surv[jl >> 5], 1u << (jl & 31));
}

__device__ __forceinline__ uint32_t df_of_fields(const uint32_t f[5])
{
    return ((f[0] & 1u) << 4) | ((f[1] & 1u) << 3) | ((f[2] & 1u) << 2) | ((f[3] & 1u) << 1) | (f[4] & 1u);
}

template <bool FROM_MAG>
__global__ void __launch_bounds__(kThreads, B200_SCAN_MIN_BLOCKS) scan_kernel(const ScanParams p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const ScanParams &L = p;   // layout fields
    uint16_t *mag = reinterpret_cast<uint16_t *>(smem);
    u64x *dd2 = reinterpret_cast<u64x *>(smem + L.off_dd);                  // (A, B) first-difference pairs
    uint32_t *fb = reinterpret_cast<uint32_t *>(smem + L.off_dd);           // P4: staged fields (dd2 is dead)
    const uint32_t *lut = p.lut;                                             // [12][5][5] field extraction table
    uint32_t *planes = reinterpret_cast<uint32_t *>(smem + L.off_planes);   // [5][12][WP]
    uint8_t *Rc = smem + L.off_edges;                 // rising-edge bit of every sample (bit i <-> m[i] < m[i+1])
    uint8_t *Fc = Rc + L.edge_bytes;                  // falling-edge bit
    uint32_t *surv = reinterpret_cast<uint32_t *>(smem + L.off_surv);
    const uint32_t *tabs = p.crc_tabs;
    uint16_t *queue = reinterpret_cast<uint16_t *>(smem + L.off_queue);     // [5][kQueueCap]
    uint16_t *cand = reinterpret_cast<uint16_t *>(smem + L.off_cand);
    __shared__ uint32_t s_warp_tot[kWarps];
    __shared__ uint32_t s_base, s_count, s_ok, s_qn[5], s_nlong, s_nshort;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint32_t b = tile / (uint32_t)p.tiles_per_buffer;
    const int k = (int)(tile - b * (uint32_t)p.tiles_per_buffer);
    const int len = p.lengths ? (int)min(p.lengths[b], p.spb) : (int)p.spb;
    const int tile_start = k * p.T;
    if (tile_start >= len) {
        if (tid == 0)
            p.tile_dir[tile] = make_uint2(0u, 0u);
        return;
    }
    const int npos = min(p.T, len - tile_start);
    const int steps = (npos + kHaloTot + kStep - 1) / kStep;   // 384-blocks actually needed
    const int WP = L.WP;

    // ---- P1/P2 in two passes over the tile (so that the first-difference buffer holds half
    // of it and four thread blocks fit one SM).
    // Pair-slot s holds sample s of stream A (tile samples [0, 192H)) and sample 192H+s of
    // stream B.  Pass `ps` covers the half blocks [hb0, hb0+nh) of both streams.
    const int H = steps;
    const int offB = kHalf * H;
    const int Hh = (H + 1) / 2;
    int prev_len;
    const uint32_t *prev = FROM_MAG ? nullptr
                                    : carry_source(p.in, p.stride, p.lengths, p.spb, b, p.carry, p.tail, &prev_len);
    for (int c = tid; c < L.nw; c += kThreads)
        surv[c] = 0;
    for (int c = tid; c < 5 * 12; c += kThreads)
        planes[c * WP + steps] = 0;   // pad word read by funnel shifts
    if (tid < 5)
        s_qn[tid] = 0;
    for (int ps = 0; ps < 2; ps++) {
        const int hb0 = ps * Hh, nh = min(Hh, H - hb0);
        if (nh <= 0)
            break;
        // ---- P1: magnitudes (u16), edge bits and first differences -> shared memory.  A warp
        // takes 248 new pair-slots per iteration, 8 per lane (2 x 2 LDG.128); lane 31 recomputes
        // the next chunk's first 8 only to hand lane 30 its right neighbour.
        {
            const int sl0 = kHalf * hb0, own_end = kHalf * (hb0 + nh);
            const int nslots = kHalf * nh + 16;   // + right neighbours / pad mirror of the last half block
            const int nchunk = (nslots + kChunk - 1) / kChunk;
            const int s0 = tile_start - (kTrailing + kHaloFront);   // sample index of m[0]
            const int i0 = tile_start - kHaloFront;                  // data index of m[0]
            const uint32_t *b32 = reinterpret_cast<const uint32_t *>(p.in) + (unsigned long long)b * p.stride;
            const uint16_t *d16 = reinterpret_cast<const uint16_t *>(p.in) + (unsigned long long)b * p.stride;
            for (int ch = warp; ch < nchunk; ch += kWarps) {
                const int rel = ch * kChunk + 8 * lane;   // pair-slot relative to this pass
                const int sl = sl0 + rel;
                u64x r[9];   // (A, B) pairs of f32 bit patterns 0x4B000000 + magnitude = 2^23 + magnitude
                if (!FROM_MAG) {
                    uint32_t wa[8], wb[8];
                    load_iq8(b32, s0 + sl, len, p.vec_ok, prev, prev_len, wa);
                    load_iq8(b32, s0 + offB + sl, len, p.vec_ok, prev, prev_len, wb);
#pragma unroll
                    for (int e = 0; e < 8; e++)
                        r[e] = mag_pair_fast2(wa[e], wb[e]);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int ia = i0 + sl + e, ib = ia + offB;
                        const uint32_t ma = (ia >= 0 && ia < kMagLen) ? (uint32_t)__ldg(d16 + ia) : 0u;
                        const uint32_t mb = (ib >= 0 && ib < kMagLen) ? (uint32_t)__ldg(d16 + ib) : 0u;
                        r[e] = f2_pack(__uint_as_float(0x4B000000u + ma), __uint_as_float(0x4B000000u + mb));
                    }
                }
                r[8] = __shfl_down_sync(0xffffffffu, r[0], 1);
                if (lane < 31 && rel < nslots) {
                    u64x dv[8];
#pragma unroll
                    for (int e = 0; e < 8; e++)
                        dv[e] = f2_sub(r[e + 1], r[e]);      // m[i+1]-m[i], exact
                    if (sl < own_end) {   // the 16 extra pair-slots belong to the next pass / tile
                        uint32_t ra[8], rb[8];
                        uint32_t fA = 0, fB = 0, rA_ = 0, rB_ = 0;
#pragma unroll
                        for (int e = 0; e < 8; e++) {
                            float x, y;
                            f2_unpack(r[e], x, y);
                            ra[e] = __float_as_uint(x);
                            rb[e] = __float_as_uint(y);
                        }
                        // edge bits of these samples (demod_2400.rs:221-317 compares neighbours only)
#pragma unroll
                        for (int e = 7; e >= 0; e--) {
                            float x, y, nx, ny;
                            f2_unpack(dv[e], x, y);
                            f2_unpack(f2_sub(r[e], r[e + 1]), nx, ny);
                            fA = __funnelshift_l(__float_as_uint(x), fA, 1);     // m[i+1]-m[i] < 0: falling
                            fB = __funnelshift_l(__float_as_uint(y), fB, 1);
                            rA_ = __funnelshift_l(__float_as_uint(nx), rA_, 1);  // rising
                            rB_ = __funnelshift_l(__float_as_uint(ny), rB_, 1);
                        }
                        *reinterpret_cast<uint4 *>(mag + sl) =
                            make_uint4(__byte_perm(ra[0], ra[1], 0x5410), __byte_perm(ra[2], ra[3], 0x5410),
                                       __byte_perm(ra[4], ra[5], 0x5410), __byte_perm(ra[6], ra[7], 0x5410));
                        *reinterpret_cast<uint4 *>(mag + offB + sl) =
                            make_uint4(__byte_perm(rb[0], rb[1], 0x5410), __byte_perm(rb[2], rb[3], 0x5410),
                                       __byte_perm(rb[4], rb[5], 0x5410), __byte_perm(rb[6], rb[7], 0x5410));
                        Fc[sl >> 3] = (uint8_t)fA;
                        Rc[sl >> 3] = (uint8_t)rA_;
                        Fc[(offB + sl) >> 3] = (uint8_t)fB;
                        Rc[(offB + sl) >> 3] = (uint8_t)rB_;
                    }
                    const int qb = rel / kQuarter, off = rel - qb * kQuarter;   // quarter block inside this pass
                    u64x *dst = dd2 + kQuarterPad * qb + off;
#pragma unroll
                    for (int e = 0; e < 8; e += 2)
                        *reinterpret_cast<ulonglong2 *>(dst + e) = make_ulonglong2(dv[e], dv[e + 1]);
                    if (off < 12 && qb > 0) {   // mirror into the previous quarter block's pad
                        u64x *pad = dd2 + kQuarterPad * (qb - 1) + kQuarter + off;
                        *reinterpret_cast<ulonglong2 *>(pad) = make_ulonglong2(dv[0], dv[1]);
                        *reinterpret_cast<ulonglong2 *>(pad + 2) = make_ulonglong2(dv[2], dv[3]);
                        if (off == 0) {
                            *reinterpret_cast<ulonglong2 *>(pad + 4) = make_ulonglong2(dv[4], dv[5]);
                            *reinterpret_cast<ulonglong2 *>(pad + 6) = make_ulonglong2(dv[6], dv[7]);
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- P2: correlator sign planes.  One lane = 8 samples at stride 12 (residue class rho,
        // steps 8*sub..8*sub+7) of half block hb of stream A and of stream B, processed as
        // (A, B) pairs in the packed f32x2 pipe; sign bits are shifted in with funnel shifts.
        // demod_2400.rs:72-83 on negated first differences u,v,w (so that "x > 0" is a sign bit):
        //   -[5,-3,-2] = 5u+2v   -[4,-1,-3] = 4u+3v   -[3,1,-4] = 3u+4v   -[2,3,-5] = 2u+5v
        //   -[1,5,-5,-1] = u+6v+w      (all exact in f32: |.| < 2^20)
        {
            const u64x five = f2_pack(5.0f, 5.0f);
            uint8_t *pbytes = reinterpret_cast<uint8_t *>(planes);
            for (int task = tid; task < 24 * nh; task += kThreads) {
                const int qb = task / 12, rho = task - 12 * qb;
                const u64x *dp = dd2 + kQuarterPad * qb + rho;
                uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 0;
#pragma unroll
                for (int q = 7; q >= 0; q--) {
                    const u64x u = dp[12 * q], v = dp[12 * q + 1], w = dp[12 * q + 2];
                    const u64x g = f2_sub(v, u);
                    const u64x x0 = f2_fma(u, five, f2_add(v, v));
                    const u64x x1 = f2_add(x0, g), x2 = f2_add(x1, g), x3 = f2_add(x2, g);
                    const u64x x4 = f2_add(f2_add(x3, g), w);
                    float lo, hi;
                    f2_unpack(x0, lo, hi);
                    a0 = __funnelshift_l(__float_as_uint(lo), a0, 1);
                    b0 = __funnelshift_l(__float_as_uint(hi), b0, 1);
                    f2_unpack(x1, lo, hi);
                    a1 = __funnelshift_l(__float_as_uint(lo), a1, 1);
                    b1 = __funnelshift_l(__float_as_uint(hi), b1, 1);
                    f2_unpack(x2, lo, hi);
                    a2 = __funnelshift_l(__float_as_uint(lo), a2, 1);
                    b2 = __funnelshift_l(__float_as_uint(hi), b2, 1);
                    f2_unpack(x3, lo, hi);
                    a3 = __funnelshift_l(__float_as_uint(lo), a3, 1);
                    b3 = __funnelshift_l(__float_as_uint(hi), b3, 1);
                    f2_unpack(x4, lo, hi);
                    a4 = __funnelshift_l(__float_as_uint(lo), a4, 1);
                    b4 = __funnelshift_l(__float_as_uint(hi), b4, 1);
                }
                // stream bit q = 8*(2*hb0 + qb) + step: one byte per task and stream
                uint8_t *pa = pbytes + 4 * (rho * WP) + 2 * hb0 + qb, *pb = pa + 2 * H;
                const int pstride = 4 * 12 * WP;
                pa[0 * pstride] = (uint8_t)a0; pb[0 * pstride] = (uint8_t)b0;
                pa[1 * pstride] = (uint8_t)a1; pb[1 * pstride] = (uint8_t)b1;
                pa[2 * pstride] = (uint8_t)a2; pb[2 * pstride] = (uint8_t)b2;
                pa[3 * pstride] = (uint8_t)a3; pb[3 * pstride] = (uint8_t)b3;
                pa[4 * pstride] = (uint8_t)a4; pb[4 * pstride] = (uint8_t)b4;
            }
        }
        __syncthreads();
    }

    // ---- P3a: preamble templates on the edge bitmaps, 32 consecutive positions per thread;
    // matches go to one queue per template case
    {
        const uint32_t *R32 = reinterpret_cast<const uint32_t *>(Rc), *F32 = reinterpret_cast<const uint32_t *>(Fc);
        const int nwp = (npos + kHaloFront + 31) / 32;
        for (int w = tid; w < nwp; w += kThreads) {
            // valid positions: 2 <= mi < npos+2 with mi = 32w + bit
            const int lo = max(kHaloFront - 32 * w, 0), hi = min(npos + kHaloFront - 32 * w, 32);
            const uint32_t valid = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
            const uint32_t r0 = R32[w], r1 = R32[w + 1], f0 = F32[w], f1 = F32[w + 1];
#define EDGE_R(s) __funnelshift_r(r0, r1, s)
#define EDGE_F(s) __funnelshift_r(f0, f1, s)
            const uint32_t quick = r0 & EDGE_F(12) & valid;   // demod_2400.rs:221
            if (!quick)
                continue;
            const uint32_t F1 = EDGE_F(1), F2 = EDGE_F(2), F3 = EDGE_F(3), F4 = EDGE_F(4), F9 = EDGE_F(9),
                           F10 = EDGE_F(10);
            const uint32_t R2 = EDGE_R(2), R3 = EDGE_R(3), R8 = EDGE_R(8), R9 = EDGE_R(9), R10 = EDGE_R(10),
                           R11 = EDGE_R(11);
#undef EDGE_R
#undef EDGE_F
            // demod_2400.rs:226-317, in order; first match wins
            const uint32_t T3 = F1 & R2 & F3 & R8 & F9 & R10;
            const uint32_t T4 = F1 & R2 & F3 & R8 & F9 & R11;
            const uint32_t T5 = F1 & R2 & F4 & R8 & F10 & R11;
            const uint32_t T6 = F1 & R3 & F4 & R9 & F10 & R11;
            const uint32_t T7 = F2 & R3 & F4 & R9 & F10 & R11;
            uint32_t any = quick & (T3 | T4 | T5 | T6 | T7);
            if (!any)
                continue;
            // case number as three bit planes: 0:T3 1:T4 2:T5 3:T6 4:T7
            const uint32_t c1 = T4 & ~T3, c2 = T5 & ~(T3 | T4), c3 = T6 & ~(T3 | T4 | T5),
                           c4 = ~(T3 | T4 | T5 | T6);
            const uint32_t b0 = c1 | c3, b1 = c2 | c3;
            while (any) {
                const int bit = __ffs(any) - 1;
                any &= any - 1;
                const uint32_t cs = ((b0 >> bit) & 1u) | (((b1 >> bit) & 1u) << 1) | (((c4 >> bit) & 1u) << 2);
                const int mi = 32 * w + bit;
                const uint32_t qi = atomicAdd(&s_qn[cs], 1u);
                if (qi < (uint32_t)kQueueCap)
                    queue[cs * kQueueCap + qi] = (uint16_t)mi;
                else
                    gate_eval(mag, surv, mi, cs);     // queue full: evaluate in place
            }
        }
    }
    __syncthreads();
    // ---- P3b: SNR and quiet-zone gates, one queue entry per thread, queues back to back
    {
        int n[5], tot = 0;
#pragma unroll
        for (int c = 0; c < 5; c++) {
            n[c] = (int)min(s_qn[c], (uint32_t)kQueueCap);
            tot += n[c];
        }
        for (int g = tid; g < tot; g += kThreads) {
            int cs = 0, idx = g;
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (cs == c && idx >= n[c]) {
                    idx -= n[c];
                    cs = c + 1;
                }
            gate_eval(mag, surv, (int)queue[cs * kQueueCap + idx], (uint32_t)cs);
        }
    }
    __syncthreads();

    // ---- P4a: count survivors, reserve pool space (positions are emitted in ascending j)
    uint32_t wv = (tid < L.nw) ? surv[tid] : 0u;   // nw <= 256 == kThreads
    int my_off;
    {
        const int cnt = __popc(wv);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        if (lane == 31)
            s_warp_tot[warp] = (uint32_t)incl;
        __syncthreads();
        my_off = incl - cnt;
        uint32_t total = 0;
#pragma unroll
        for (int wi = 0; wi < kWarps; wi++) {
            const uint32_t t = s_warp_tot[wi];
            if (wi < warp)
                my_off += (int)t;
            total += t;
        }
        if (tid == 0) {
            uint32_t base = 0, ok = 1;
            if (total) {
                base = atomicAdd(&p.counters[C_POOL], total);
                if (base + total > p.pool_cap || base + total < base) {
                    atomicOr(&p.counters[C_FLAGS], F_POOL_OVF);
                    ok = 0;
                }
                atomicAdd(&p.counters[C_CAND], total);
            }
            p.tile_dir[tile] = make_uint2(base, ok ? total : 0u);
            s_base = base;
            s_count = total;
            s_ok = ok;
            s_nlong = 0;
            s_nshort = 0;
        }
    }
    __syncthreads();
    if (!s_ok || s_count == 0)
        return;

    // ---- P4b: five try-phases per survivor, in windows of kCandCap survivors.
    //   A: pull the five 23-bit fields of each (survivor, try_phase) out of the sign planes,
    //      read the DF; items that need a CRC are staged by class (112-bit / 56-bit syndrome)
    //   B: CRC-24 + classification on class-homogeneous runs -> record words + ICAO add-events
    const int C = (int)s_count;
    const unsigned long long ord_buf = (p.ord_first + (unsigned long long)b * p.ord_stride) << 20;
    for (int win = 0; win < C; win += kCandCap) {
        {
            uint32_t wv2 = wv;
            int off = my_off;
            while (wv2) {
                const int bit = __ffs(wv2) - 1;
                wv2 &= wv2 - 1;
                if (off >= win && off < win + kCandCap)
                    cand[off - win] = (uint16_t)(tid * 32 + bit);
                off++;
            }
        }
        __syncthreads();
        const int Cw = min(kCandCap, C - win);
        uint32_t *rec_w = p.rec + 6ull * (s_base + (uint32_t)win);
        for (int item = tid; item < 5 * Cw; item += kThreads) {
            const int c = item / 5, tt = item - 5 * c;
            const int jl = cand[c];
            // demod_2400.rs:158-160: P0 = 5*(mi+19) + try_phase, try_phase = 4+tt
            const int A = jl + kHaloFront + 19;
            const int qA = A / 12, rA = A - 12 * qA;
            const uint32_t *lrow = lut + 25 * rA + 5 * tt;
            uint32_t f[5];
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const uint32_t e = __ldg(lrow + r);
                const int q = qA + (int)(e >> 16);
                const uint32_t *st = planes + (e & 0xffffu) + (q >> 5);
                f[r] = __funnelshift_r(st[0], st[1], q & 31) & (r < 2 ? 0x7fffffu : 0x3fffffu);
            }
            if (tt == 0)
                rec_w[6 * c] = (uint32_t)(tile_start + jl);
            // class of the item: 0 = decided here, 1 = needs the 112-bit syndrome, 2 = the 56-bit one
            uint32_t wd = 0;
            int cls = 0;
            if ((f[0] | f[1] | f[2] | f[3] | f[4]) == 0) {
                wd = kNoneMarker;                      // all 14 bytes zero -> None (mode_s/mod.rs:51-53)
            } else {
                const uint32_t bit = 1u << df_of_fields(f);
                if (bit & 0xFF370000u)                 // DF 16,17,18,20,21,24..31
                    cls = 1;
                else if (bit & 0x00000831u)            // DF 0,4,5,11
                    cls = 2;
            }
#if B200_AGG_ATOMICS
            // one shared-memory atomic per warp and class instead of one per item
            const unsigned act = __activemask();
            const unsigned ml = __ballot_sync(act, cls == 1), ms = __ballot_sync(act, cls == 2);
            const int leader = __ffs(act) - 1;
            uint32_t basel = 0, bases = 0;
            if (lane == leader) {
                if (ml)
                    basel = atomicAdd(&s_nlong, (uint32_t)__popc(ml));
                if (ms)
                    bases = atomicAdd(&s_nshort, (uint32_t)__popc(ms));
            }
            basel = __shfl_sync(act, basel, leader);
            bases = __shfl_sync(act, bases, leader);
            const unsigned lt = (1u << lane) - 1u;
#endif
            if (cls) {
#if B200_AGG_ATOMICS
                const uint32_t slot = cls == 1 ? basel + (uint32_t)__popc(ml & lt)
                                               : (uint32_t)(kFieldItems - 1) - (bases + (uint32_t)__popc(ms & lt));
#else
                const uint32_t slot = cls == 1 ? atomicAdd(&s_nlong, 1u)
                                               : (uint32_t)(kFieldItems - 1) - atomicAdd(&s_nshort, 1u);
#endif
                uint32_t *o = fb + 5 * slot;
                o[0] = f[0];
                o[1] = f[1];
                o[2] = f[2] | (((uint32_t)item & 0x3ffu) << 22);
                o[3] = f[3] | (((uint32_t)item >> 10) << 22);
                o[4] = f[4];
            } else {
                rec_w[6 * c + 1 + tt] = wd;
            }
        }
        __syncthreads();
        {
            const int nl = (int)s_nlong, ns = (int)s_nshort;
            for (int g = tid; g < nl + ns; g += kThreads) {
                const bool is_long = g < nl;
                const uint32_t slot = is_long ? (uint32_t)g : (uint32_t)(kFieldItems - 1 - (g - nl));
                const uint32_t *o = fb + 5 * slot;
                uint32_t f[5] = {o[0], o[1], o[2], o[3], o[4]};
                const int item = (int)((f[2] >> 22) | ((f[3] >> 22) << 10));
                f[2] &= 0x3fffffu;
                f[3] &= 0x3fffffu;
                const uint32_t df = df_of_fields(f);
                uint32_t wd;
                if (is_long) {
                    const uint32_t syn = syn112_fields(tabs, f);
                    if (df == 17 || df == 18)          // mode_s/mod.rs:91-109
                        wd = syn ? 0u : (((df == 17 ? K_DF17 : K_DF18) << 29) | msg_bits<8, 24>(f));
                    else                                // :110-134
                        wd = (K_PAR_LONG << 29) | syn;
                } else {
                    const uint32_t syn = syn56_fields(tabs, f);
                    if (df == 11)                       // :73-90
                        wd = (syn & 0xffff80u) ? 0u
                                               : ((((syn & 0x7f) ? K_DF11_IID : K_DF11_IID0) << 29) | msg_bits<8, 24>(f));
                    else                                // :56-72
                        wd = (K_PAR_SHORT << 29) | syn;
                }
                const int c = item / 5, tt = item - 5 * c;
                rec_w[6 * c + 1 + tt] = wd;
                const uint32_t kind = wd >> 29;
                if (kind == K_DF11_IID0 || kind == K_DF17 || kind == K_DF18) {
                    const uint32_t key = (wd & 0xffffffu) | (kind == K_DF18 ? B200ADSB_ICAO_FILTER_ADSB_NT : 0u);
                    const uint32_t j = (uint32_t)(tile_start + cand[c]);
                    event_add(p.ev_keys, p.ev_ord, p.ev_used, p.ev_mask, p.counters, key,
                              ord_buf | ((unsigned long long)j << 3) | (unsigned long long)tt);
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            s_nlong = 0;
            s_nshort = 0;
        }
    }
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
// packed form for a sync-free exchange: rows[0] = (count, 0), rows[1..] = (key, ordinal)
__global__ void events_pack_kernel(const uint32_t *ev_keys, const unsigned long long *ev_ord,
                                   const uint32_t *ev_used, const uint32_t *counters,
                                   unsigned long long *rows, uint32_t rows_cap)
{
    const uint32_t n = counters[C_EV_USED];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        rows[0] = n;
        rows[1] = 0;
    }
    const uint32_t m = min(n, rows_cap - 1);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        const uint32_t h = ev_used[i];
        rows[2ull * (i + 1)] = ev_keys[h];
        rows[2ull * (i + 1) + 1] = ev_ord[h];
    }
}
// gathered: n_ranks blocks of rows_per_rank packed rows; merges every block but `skip`
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
        if (n > rows_per_rank - 1) {   // that rank had more events than fit the exchange buffer
            if (blockIdx.x == 0 && threadIdx.x == 0)
                atomicOr(&counters[C_FLAGS], F_EV_OVF);
            continue;
        }
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)n; i += gridDim.x * blockDim.x)
            event_add(ev_keys, ev_ord, ev_used, mask, counters, (uint32_t)rows[2ull * (i + 1)], rows[2ull * (i + 1) + 1]);
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

// after resolve: the admitted keys join the filter; the event table is recycled
__device__ __forceinline__ void events_commit_body(uint32_t *ev_keys, unsigned long long *ev_ord,
                                                   const uint32_t *ev_used, const uint32_t *new_keys,
                                                   uint32_t *counters, uint32_t *members)
{
    // a scan that overflowed the candidate pool or the event table is redone by the host:
    // nothing of it may reach the filter
    const bool bad = (counters[C_FLAGS] & (F_POOL_OVF | F_EV_OVF)) != 0;
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
    }
}
// d_result (nullable): the enqueue-only batch outcome {frames, overflow flags, candidates,
// frames > cap}; the per-batch counters are then cleared here instead of by host memsets
__global__ void __launch_bounds__(1024) events_commit_kernel(uint32_t *ev_keys, unsigned long long *ev_ord,
                                                             const uint32_t *ev_used, const uint32_t *new_keys,
                                                             uint32_t *counters, uint32_t *members,
                                                             uint32_t *d_result = nullptr, uint32_t cap = 0)
{
    events_commit_body(ev_keys, ev_ord, ev_used, new_keys, counters, members);
    if (d_result && threadIdx.x == 0) {
        d_result[0] = counters[C_FRAMES];
        d_result[1] = counters[C_FLAGS] & (F_POOL_OVF | F_EV_OVF);
        d_result[2] = counters[C_CAND];
        d_result[3] = counters[C_FRAMES] > cap ? 1u : 0u;
        counters[C_POOL] = 0;
        counters[C_FLAGS] = 0;
        counters[C_CAND] = 0;
    }
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
    const uint32_t *tail;
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
    int prev_len = 0;
    const uint32_t *prev = (FROM_MAG || p.msgs) ? nullptr
                                                : carry_source(p.in, p.stride, p.lengths, p.spb, b, p.carry, p.tail, &prev_len);
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
                        m = mag_bits_fast(iq_word(bb, s, len, prev, prev_len)) & 0xffffu;
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
struct CommitArgs {
    uint32_t *ev_keys;
    unsigned long long *ev_ord;
    const uint32_t *ev_used, *new_keys;
    uint32_t *counters, *members, *d_result;
    uint32_t cap;
};
__device__ __forceinline__ void commit_tail(const CommitArgs &a)
{
    events_commit_body(a.ev_keys, a.ev_ord, a.ev_used, a.new_keys, a.counters, a.members);
    if (a.d_result && threadIdx.x == 0) {      // enqueue-only batch outcome, per-batch counters cleared
        a.d_result[0] = a.counters[C_FRAMES];
        a.d_result[1] = a.counters[C_FLAGS] & (F_POOL_OVF | F_EV_OVF);
        a.d_result[2] = a.counters[C_CAND];
        a.d_result[3] = a.counters[C_FRAMES] > a.cap ? 1u : 0u;
        a.counters[C_POOL] = 0;
        a.counters[C_FLAGS] = 0;
        a.counters[C_CAND] = 0;
    }
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
        int prev_len = 0;
        const uint32_t *prev = (FROM_MAG || p.msgs) ? nullptr
                                                    : carry_source(p.in, p.stride, p.lengths, p.spb, b, p.carry, p.tail, &prev_len);
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
                    m = mag_bits_fast(iq_word(bb, idx - kTrailing, len, prev, prev_len)) & 0xffffu;   // == mag_pair
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
                                                                        const EmitParams ep, const uint32_t n_ctas)
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
    events_commit_body(fa.ev_keys, fa.ev_ord, fa.ev_used, fa.new_keys, fa.counters, fa.members);
}

// carry mode: the last 326 samples of the stream (walking back over this batch's buffers,
// then into the old tail) become the next batch's leading samples
__global__ void save_tail_kernel(const uint32_t *in, unsigned long long stride, const uint32_t *lengths,
                                 uint32_t spb, uint32_t n_buffers, const uint32_t *old_tail, uint32_t *new_tail)
{
    const int k = threadIdx.x;
    if (k >= kTrailing)
        return;
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
