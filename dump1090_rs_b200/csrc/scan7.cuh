// dump1090_rs_b200/csrc/scan7.cuh -- the stage-1 kernel (sm_100a): one thread block per tile of T output
// positions -> 24-byte records {j, five classified try-phases} of the positions that pass the preamble
// gates + ICAO add-events.  Magnitudes, first differences, correlator signs and edge bits never leave
// registers between the IQ load and the bit planes.
//
// Dense phase.  The halo-extended tile is NG groups of 192 samples.  Three lanes own a group:
// lane c (0..2) takes, for slot k = 15..0, the four consecutive samples 192G + 12k + 4c + e
// (one 16-byte IQ row, brought in by cp.async through a per-lane ring).  Residue rho = 4c + e of
// the 12-sample Mode-S bit period is therefore fixed per (lane, e) and slot k is bit k of a
// 16-bit accumulator: after 16 slots every accumulator IS one halfword of the mod-12
// de-interleaved plane[f][rho] (bit q <-> tile magnitude index 12q + rho), stored with one
// STS.U16.  The three magnitudes to the right of a row come from lane+1 (c < 2) or from
// lane-2's previous slot (c == 2) by shuffle.  Samples are paired (e, e+2) in the packed f32x2
// pipe: the pairs (d0,d2) (d1,d3) (d2,d4) (d3,d5) of first differences serve both correlator pairs.
//   src/utils.rs:43-58 (magnitude), src/demod_2400.rs:62-83 (correlators), :221-317 (edges)
//
// Sparse phases.
//   P3a  one warp, lane = word column, residue static: the five templates as AND of the
//        rising/falling planes (32 positions at stride 12 per word) -> per (rho, word) the match
//        mask and the template case as three bit planes
//   P3b  warps expand 32 mask words per round into one (word, bit) list (warp scan + one shared
//        atomic per round; order is irrelevant because the gates only set survivor bits); the list
//        lives in the edge planes, which are dead by then
//   P3c  SNR and quiet-zone gates, one match per thread, no branch per template case
//                                                                    (src/demod_2400.rs:129-146)
//   P4   one warp scans the survivor bitmap and reserves pool space; all warps list the survivors;
//        per (survivor, try-phase): five 23-bit field extracts from the planes, DF, items that need a
//        CRC staged by class; CRC-24 by shuffle-table field sums; record words; ICAO add-events
//                                                    (src/mode_s/mod.rs:34-139, src/crc.rs:263-282)
//
// The kernel is compiled three times: generic (any tile size, lengths, alignment, carry), with the
// default tile as a template constant (the shared memory plan becomes immediates) and, on top of that,
// for batches of whole standard buffers.  What was measured and dropped is listed in profiles/README.md;
// -DB200_SCAN7_STOP / -DB200_ABL give the measurement builds used there.
#pragma once
#include "kernels.cuh"

namespace b200 {

constexpr int k7Threads = 128;
constexpr int k7Warps = k7Threads / 32;
constexpr int k7Slots = 16;
constexpr int k7Group = 12 * k7Slots;       // 192 samples
constexpr int k7GroupsPerWarp = 10;         // 30 of 32 lanes
constexpr int k7GroupsPerPass = k7GroupsPerWarp * k7Warps;
constexpr int k7CandCap = 128;              // survivors decoded per window
constexpr int k7FieldItems = 5 * k7CandCap;
constexpr int k7ListCap = 1024;             // template matches gated per round (at most; see Scan7Smem::list_cap)
#ifndef B200_SCAN7_MIN_BLOCKS
#define B200_SCAN7_MIN_BLOCKS 7
#endif
#ifndef B200_SCAN7_RING
#define B200_SCAN7_RING 3                   // IQ rows in flight per lane (cp.async ring)
#endif
#ifndef B200_SCAN7_STOP
#define B200_SCAN7_STOP 0                   // measurement builds only: 1 = return after the dense phase, 2 = after the gates
#endif
#ifndef B200_SCAN7_STAGGER
#define B200_SCAN7_STAGGER 1500             // ns of start delay per residency rank of a first-wave block (0: off)
#endif
#ifndef B200_ABL
#define B200_ABL 0                          // measurement builds only (with B200_SCAN7_STOP=1): ablations of the dense loop, results are wrong
#endif                                      //   1 no global loads  2 no neighbour shuffles  4 signs by FADD  16 no magnitude arithmetic  32 no magnitude store

// shared memory plan for tile size T
struct Scan7Smem {
    int NG, WP, nw, mag_len, list_cap;
    size_t off_planes, off_surv, off_masks, off_list, off_cand, bytes;
    __host__ __device__ constexpr explicit Scan7Smem(int T)
        : NG(0), WP(0), nw(0), mag_len(0), list_cap(0), off_planes(0), off_surv(0), off_masks(0), off_list(0),
          off_cand(0), bytes(0)
    {
        NG = (T + kHaloTot + k7Group - 1) / k7Group;
        mag_len = NG * k7Group + 32;
        WP = (NG + 1) / 2 + 1;                        // words per plane row (+1 read by funnel shifts)
        nw = (T + 31) / 32;
        size_t o = (size_t)mag_len * 2;
        if (o < (size_t)k7FieldItems * 5 * 4)         // P4 field staging aliases the (dead) magnitudes
            o = (size_t)k7FieldItems * 5 * 4;
        o = (o + 15) & ~(size_t)15;
        off_planes = o;                               // [7][12][WP]: 5 correlators, rising, falling
        o += (size_t)7 * 12 * WP * 4;
        off_surv = o;
        o += (size_t)nw * 4;
        o = (o + 15) & ~(size_t)15;
        off_masks = o;                                // per-lane ring of IQ rows in flight (cp.async), 16 B each;
        {                                             // later the template match masks [12][WP] x 16 B
            const size_t ring_b = (size_t)k7Threads * B200_SCAN7_RING * 16, mask_b = (size_t)12 * WP * 16;
            o += ring_b > mask_b ? ring_b : mask_b;
        }
        // the match list lives in the rising + falling planes (2 * 12 * WP words), dead after the template pass
        off_list = off_planes + (size_t)5 * 12 * WP * 4;
        list_cap = 48 * WP < k7ListCap ? 48 * WP : k7ListCap;
        off_cand = o;
        o += (size_t)k7CandCap * 2;
        bytes = (o + 15) & ~(size_t)15;
    }
};

struct Scan7Params {
    ScanParams s;
    uint32_t off_planes, off_surv, off_masks, off_list, off_cand;
    int NG, WP, nw, list_cap;
    int Wrow, inv_Wrow;    // mask words per residue row for this tile size; ceil(65536 / Wrow): i / Wrow == (i * inv) >> 16 for i < 12 * Wrow
};

// one row of four consecutive magnitudes, as f32 bit patterns 2^23 + m, paired (m0,m2) (m1,m3)
struct Row {
    u64x q0, q1;
};

template <bool FROM_MAG>
__device__ __noinline__ Row row_slow(const ScanParams &p, const uint32_t *b32, const uint16_t *d16, int s,
                                        int i, int len, const CarrySrc &cs, uint32_t b)
{
    Row r;
    if (!FROM_MAG) {
        uint32_t w[4];
        if (s >= 0 && s + 3 < len && p.vec_ok) {
            const int4 v = __ldg(reinterpret_cast<const int4 *>(b32 + s));
            w[0] = (uint32_t)v.x; w[1] = (uint32_t)v.y; w[2] = (uint32_t)v.z; w[3] = (uint32_t)v.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++)
                w[e] = iq_word(b32, s + e, len, cs, b);
        }
        r.q0 = mag_pair_fast2(w[0], w[2]);
        r.q1 = mag_pair_fast2(w[1], w[3]);
    } else {
        uint32_t m[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int ia = i + e;
            m[e] = 0x4B000000u + ((ia >= 0 && ia < kMagLen) ? (uint32_t)__ldg(d16 + ia) : 0u);
        }
        r.q0 = f2_pack(__uint_as_float(m[0]), __uint_as_float(m[2]));
        r.q1 = f2_pack(__uint_as_float(m[1]), __uint_as_float(m[3]));
    }
    return r;
}

__device__ __forceinline__ uint32_t sgn_in(float x, uint32_t acc)   // acc = acc << 1 | sign(x)
{
#if B200_ABL & 4
    return __float_as_uint(__fadd_rn(__uint_as_float(acc), x));
#else
    return __funnelshift_l(__float_as_uint(x), acc, 1);
#endif
}
__device__ __forceinline__ u64x mag_pair_abl(uint32_t wa, uint32_t wb)
{
#if B200_ABL & 16
    return f2_pack(__uint_as_float(0x4B000000u | (wa & 0xffffu)), __uint_as_float(0x4B000000u | (wb & 0xffffu)));
#else
    return mag_pair_fast2(wa, wb);
#endif
}

// the five PPM correlators of a sample pair on first differences u, v, w
// (demod_2400.rs:72-83 negated so that "bit = 1" is the sign bit):
//   5u+2v, 4u+3v, 3u+4v, 2u+5v, u+6v+w      all exact in f32 (|.| < 2^20)
__device__ __forceinline__ void corr_pair(u64x u, u64x v, u64x w, uint32_t *lo, uint32_t *hi)
{
    const u64x five = f2_pack(5.0f, 5.0f);
    const u64x g = f2_sub(v, u);
    const u64x x0 = f2_fma(u, five, f2_add(v, v));
    const u64x x1 = f2_add(x0, g), x2 = f2_add(x1, g), x3 = f2_add(x2, g);
    const u64x x4 = f2_add(f2_add(x3, g), w);
    float a, b;
    f2_unpack(x0, a, b); lo[0] = sgn_in(a, lo[0]); hi[0] = sgn_in(b, hi[0]);
    f2_unpack(x1, a, b); lo[1] = sgn_in(a, lo[1]); hi[1] = sgn_in(b, hi[1]);
    f2_unpack(x2, a, b); lo[2] = sgn_in(a, lo[2]); hi[2] = sgn_in(b, hi[2]);
    f2_unpack(x3, a, b); lo[3] = sgn_in(a, lo[3]); hi[3] = sgn_in(b, hi[3]);
    f2_unpack(x4, a, b); lo[4] = sgn_in(a, lo[4]); hi[4] = sgn_in(b, hi[4]);
}

struct DenseState {
    uint32_t acc[4][7];      // [e][plane]: planes 0..4 correlators, 5 rising, 6 falling
    float pm0, pm1, pm2;     // first three magnitudes of the row 12 samples to the right (c == 0 lanes)
};

// one slot: row r -> magnitudes to shared memory, 28 sign bits into the accumulators
__device__ __forceinline__ void dense_slot(DenseState &st, const Row r, bool is_c0, int src_lane, uint16_t *mag_row,
                                           bool store)
{
    float m0, m1, m2, m3;
    f2_unpack(r.q0, m0, m2);
    f2_unpack(r.q1, m1, m3);
    // right neighbours m4..m6: lane+1's m0..m2 of this slot, or (c == 2) lane-2's m0..m2 of the
    // previous slot = the row 12 samples further
    const float t0 = is_c0 ? st.pm0 : m0, t1 = is_c0 ? st.pm1 : m1, t2 = is_c0 ? st.pm2 : m2;
#if B200_ABL & 2
    const float m4 = t0, m5 = t1, m6 = t2;
#else
    const float m4 = __shfl_sync(0xffffffffu, t0, src_lane);
    const float m5 = __shfl_sync(0xffffffffu, t1, src_lane);
    const float m6 = __shfl_sync(0xffffffffu, t2, src_lane);
#endif
    st.pm0 = m0; st.pm1 = m1; st.pm2 = m2;
    // first differences as pairs (d0,d2) (d1,d3) (d2,d4) (d3,d5); only the first comes out of aligned
    // register pairs, the others are cheaper as scalar subtractions written straight into their pair
    // than as packed ones fed by register moves
    const u64x p0 = f2_sub(r.q1, r.q0), n0 = f2_sub(r.q0, r.q1);
    const u64x p1 = f2_pack(__fsub_rn(m2, m1), __fsub_rn(m4, m3));
    const u64x p2 = f2_pack(__fsub_rn(m3, m2), __fsub_rn(m5, m4));
    const u64x p3 = f2_pack(__fsub_rn(m4, m3), __fsub_rn(m6, m5));
    const u64x n1 = f2_pack(__fsub_rn(m1, m2), __fsub_rn(m3, m4));
    float a, b;
    f2_unpack(p0, a, b); st.acc[0][6] = sgn_in(a, st.acc[0][6]); st.acc[2][6] = sgn_in(b, st.acc[2][6]);   // falling
    f2_unpack(p1, a, b); st.acc[1][6] = sgn_in(a, st.acc[1][6]); st.acc[3][6] = sgn_in(b, st.acc[3][6]);
    f2_unpack(n0, a, b); st.acc[0][5] = sgn_in(a, st.acc[0][5]); st.acc[2][5] = sgn_in(b, st.acc[2][5]);   // rising
    f2_unpack(n1, a, b); st.acc[1][5] = sgn_in(a, st.acc[1][5]); st.acc[3][5] = sgn_in(b, st.acc[3][5]);
    corr_pair(p0, p1, p2, st.acc[0], st.acc[2]);
    corr_pair(p1, p2, p3, st.acc[1], st.acc[3]);
    if (store && !(B200_ABL & 32)) {
        const uint2 v = make_uint2(__byte_perm(__float_as_uint(m0), __float_as_uint(m1), 0x5410),
                                   __byte_perm(__float_as_uint(m2), __float_as_uint(m3), 0x5410));
        *reinterpret_cast<uint2 *>(mag_row) = v;
    }
}

__device__ __forceinline__ void gate_eval7(const uint16_t *mag, uint32_t *surv, int mi, uint32_t cs, int npos)
{
    const int jl = mi - kHaloFront;
    if (jl < 0 || jl >= npos)
        return;
    gate_eval(mag + mi, surv, jl, cs);
}
__device__ __noinline__ void gate_eval7_cold(const uint16_t *mag, uint32_t *surv, int mi, uint32_t cs, int npos)
{
    gate_eval7(mag, surv, mi, cs, npos);
}

// SNR and quiet-zone gates of one template match (demod_2400.rs:129,135-146) without a branch
// per template case.  With a=p1 h=p2 b=p3 e=p4 c=p9 f=p10 g=p11 d=p12 (demod_2400.rs:226-317):
//   case   high*4            signal        noise
//   0 (3)  a+b+c+g+d         a+b+c         p5+p6+p7
//   1 (4)  a+b+c+d           a+b+c+d       p5+p6+p7+p8
//   2 (5)  a+b+e+c+f+d       a+d           p6+p7
//   3 (6)  a+e+f+d           a+e+f+d       p5+p6+p7+p8
//   4 (7)  a+h+e+f+d         e+f+d         p6+p7+p8
__device__ __forceinline__ void gate_eval_bf(const uint16_t *mag, uint32_t *surv, int mi, uint32_t cs, int npos)
{
    const int jl = mi - kHaloFront;
    if (jl < 0 || jl >= npos)
        return;
    const uint16_t *pp = mag + mi;
    const int a = pp[1], h = pp[2], b = pp[3], e = pp[4], n5 = pp[5], n6 = pp[6], n7 = pp[7], n8 = pp[8];
    const int c = pp[9], f = pp[10], g = pp[11], d = pp[12];
    const int bc = b + c, ef = e + f;
    const int H = a + d + (cs < 3 ? bc : 0) + (cs >= 2 ? ef : 0) + (cs == 0 ? g : 0) + (cs == 4 ? h : 0);
    const int S = (cs < 4 ? a : 0) + (cs < 2 ? bc : 0) + (cs >= 1 ? d : 0) + (cs >= 3 ? ef : 0);
    const int N = n6 + n7 + (((0x0Bu >> cs) & 1u) ? n5 : 0) + (((0x1Au >> cs) & 1u) ? n8 : 0);
    if (2 * S < 3 * N)            // demod_2400.rs:129
        return;
    const int mx = max(max(max(n5, n6), max(n7, n8)),
                       max(max(max((int)pp[14], (int)pp[15]), max((int)pp[16], (int)pp[17])), (int)pp[18]));
    if (mx >= (H >> 2))           // demod_2400.rs:135-146 (high = sum / 4, non-negative)
        return;
    atomicOr(&surv[jl >> 5], 1u << (jl & 31));
}

// P3a for the residues RHO0..RHO0+NRHO-1, lane = word column w: the 32 positions 12*(32w+bit)+rho.
// X[i] is the plane word for edge offset s = i - (rho - RHO0): row t = RHO0 + i mod 12, shifted by
// one bit when t >= 12 (the carry into the next 12-sample period).  Output per (rho, w): the match
// mask and the template case as three bit planes.
template <int RHO0, int NRHO>
__device__ __forceinline__ void p3a_rows(const uint32_t *planes, int WP, int nwq, int lane, uint32_t *masks, int Wrow)
{
    const uint32_t *Rp = planes + 5 * 12 * WP, *Fp = planes + 6 * 12 * WP;
    for (int w = lane; w < nwq; w += 32) {
        uint32_t XR[NRHO + 12], XF[NRHO + 12];
#pragma unroll
        for (int i = 0; i < NRHO + 12; i++) {
            const int t = RHO0 + i;
            if (t < 12) {
                XR[i] = Rp[t * WP + w];
                XF[i] = Fp[t * WP + w];
            } else {
                XR[i] = __funnelshift_r(Rp[(t - 12) * WP + w], Rp[(t - 12) * WP + w + 1], 1);
                XF[i] = __funnelshift_r(Fp[(t - 12) * WP + w], Fp[(t - 12) * WP + w + 1], 1);
            }
        }
        uint4 *mout = reinterpret_cast<uint4 *>(masks) + w;
#pragma unroll
        for (int dr = 0; dr < NRHO; dr++) {
#define ER(s) XR[dr + (s)]
#define EF(s) XF[dr + (s)]
            const uint32_t quick = ER(0) & EF(12);   // p0 < p1 && p12 > p13 (:221)
            const uint32_t T3 = EF(1) & ER(2) & EF(3) & ER(8) & EF(9) & ER(10);
            const uint32_t T4 = EF(1) & ER(2) & EF(3) & ER(8) & EF(9) & ER(11);
            const uint32_t T5 = EF(1) & ER(2) & EF(4) & ER(8) & EF(10) & ER(11);
            const uint32_t T6 = EF(1) & ER(3) & EF(4) & ER(9) & EF(10) & ER(11);
            const uint32_t T7 = EF(2) & ER(3) & EF(4) & ER(9) & EF(10) & ER(11);
#undef ER
#undef EF
            // first match wins (:226-317)
            const uint32_t c1 = T4 & ~T3, c2 = T5 & ~(T3 | T4), c3 = T6 & ~(T3 | T4 | T5);
            mout[(RHO0 + dr) * Wrow] = make_uint4(quick & (T3 | T4 | T5 | T6 | T7), c1 | c3, c2 | c3, ~(T3 | T4 | T5 | T6));
        }
    }
}

// TC: tile size fixed at compile time (0: taken from the parameters).  For the default tile the whole shared
// memory plan -- plane row stride, offsets, list capacity, mask row width -- becomes immediates (-7 % static
// instructions, -4 % time).  STD: a batch of whole standard buffers (131,072 samples each, no per-buffer
// lengths, 16-byte aligned, no carry): buffer length, tiles per buffer and the alignment test fold as well.
template <bool FROM_MAG, int TC, bool STD>
__global__ void __launch_bounds__(k7Threads, B200_SCAN7_MIN_BLOCKS) scan7_kernel(const Scan7Params Pin)
{
    static_assert(!STD || (TC != 0 && !FROM_MAG), "the standard-batch form fixes the tile as well");
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr Scan7Smem LC(TC ? TC : 8);
    constexpr int kStdTpb = TC ? (kMaxSamples + TC - 1) / TC : 1;
    struct {
        const ScanParams &s;
        uint32_t off_planes, off_surv, off_masks, off_list, off_cand;
        int WP, nw, list_cap, Wrow, inv_Wrow;
    } P = {Pin.s,
           TC ? (uint32_t)LC.off_planes : Pin.off_planes, TC ? (uint32_t)LC.off_surv : Pin.off_surv,
           TC ? (uint32_t)LC.off_masks : Pin.off_masks, TC ? (uint32_t)LC.off_list : Pin.off_list,
           TC ? (uint32_t)LC.off_cand : Pin.off_cand, TC ? LC.WP : Pin.WP, TC ? LC.nw : Pin.nw,
           TC ? LC.list_cap : Pin.list_cap, TC ? (LC.NG + 1) / 2 : Pin.Wrow,
           TC ? (65536 + (LC.NG + 1) / 2 - 1) / ((LC.NG + 1) / 2) : Pin.inv_Wrow};
    const ScanParams &p = P.s;
    const int Tc = TC ? TC : p.T;
    const int tpb = STD ? kStdTpb : p.tiles_per_buffer;
    const int vec_ok = STD ? 1 : p.vec_ok;
    uint16_t *mag = reinterpret_cast<uint16_t *>(smem);
    uint32_t *fb = reinterpret_cast<uint32_t *>(smem);                      // P4: staged fields (mag is dead)
    uint32_t *planes = reinterpret_cast<uint32_t *>(smem + P.off_planes);   // [7][12][WP]
    uint32_t *surv = reinterpret_cast<uint32_t *>(smem + P.off_surv);
    unsigned char *ring = smem + P.off_masks;                                // [B200_SCAN7_RING][k7Threads] x 16 B
    uint32_t *masks = reinterpret_cast<uint32_t *>(smem + P.off_masks);     // P3: [12][nwq] x (match, case planes); the ring is dead
    uint16_t *list = reinterpret_cast<uint16_t *>(smem + P.off_list);
    uint16_t *cand = reinterpret_cast<uint16_t *>(smem + P.off_cand);
    const uint32_t *lut = p.lut;
    __shared__ uint32_t s_base, s_count, s_ok, s_nlong, s_nshort;
    __shared__ uint16_t s_pair_off[k7Threads];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#if B200_SCAN7_STAGGER
    // The blocks of the first residency wave start together and, tiles being equal work, stay in the same
    // phase for many tiles: all dense (FP pipes contended) or all sparse (latency bound) at once.  A start
    // offset per residency rank mixes the phases on every SM from the first tile on: +1.5 % at 1000
    // buffers, +1.8 % at 8192 (profiles/r2).
    if (blockIdx.x < 148u * B200_SCAN7_MIN_BLOCKS)
        __nanosleep((blockIdx.x / 148u) * B200_SCAN7_STAGGER);
#endif
    const uint32_t tile = p.b0 * (uint32_t)tpb + blockIdx.x;   // tile of the batch
    const uint32_t b = tile / (uint32_t)tpb;
    const int kt = (int)(tile - b * (uint32_t)tpb);
    const int len = STD ? kMaxSamples : (p.lengths ? (int)min(p.lengths[b], p.spb) : (int)p.spb);
    const int tile_start = kt * Tc;
    if (tile_start >= len) {
        if (tid == 0)
            p.tile_dir[tile] = make_uint2(0u, 0u);
        return;
    }
    const int npos = min(Tc, len - tile_start);
    const int NG = (npos + kHaloTot + k7Group - 1) / k7Group;   // groups actually needed
    const int WP = P.WP;
    const int nwq = (NG + 1) / 2;                                // plane words that carry data

    // nothing here conflicts with what the dense phase stores (different addresses), so no barrier
    // is needed before it: the survivor bitmap, the pad word read by funnel shifts and, for an odd
    // number of groups, the unowned upper halfword of each row's last data word
    for (int i = tid; i < P.nw; i += k7Threads)
        surv[i] = 0;
    for (int i = tid; i < 7 * 12; i += k7Threads) {
        planes[i * WP + nwq] = 0;
        if (NG & 1)
            reinterpret_cast<uint16_t *>(planes)[2 * (i * WP) + NG] = 0;
    }

    // ---- P1: dense phase (magnitudes, edges, correlator signs; see the header)
    {
        const CarrySrc cs{reinterpret_cast<const uint32_t *>(p.in), p.stride, p.lengths, p.spb, p.tails, p.counters,
                          (FROM_MAG || STD) ? 0 : p.carry};
        const uint32_t *b32 = reinterpret_cast<const uint32_t *>(p.in) + (unsigned long long)b * p.stride;
        const uint16_t *d16 = reinterpret_cast<const uint16_t *>(p.in) + (unsigned long long)b * p.stride;
        const int s0 = tile_start - (kTrailing + kHaloFront);    // sample index of tile magnitude 0
        const int i0 = tile_start - kHaloFront;                   // data index of tile magnitude 0
        const int gl = lane / 3, c = lane - 3 * gl;
        const int src_lane = (c < 2) ? min(lane + 1, 31) : lane - 2;
        const bool is_c0 = c == 0;
        uint16_t *planes16 = reinterpret_cast<uint16_t *>(planes);
        for (int gbase = warp * k7GroupsPerWarp; gbase < NG; gbase += k7GroupsPerPass) {
            const int G = gbase + gl;                              // lanes 30, 31: helpers without stores
            const bool own = lane < 30 && G < NG;
            const int r0 = k7Group * G + 4 * c;                    // magnitude index of slot 0
            DenseState st;
#pragma unroll
            for (int e = 0; e < 4; e++)
#pragma unroll
                for (int f = 0; f < 7; f++)
                    st.acc[e][f] = 0;
            // whole warp inside the buffer and 16-byte aligned: plain LDG.128 with one slot of prefetch
            const int g_lo = k7Group * gbase, g_hi = k7Group * (gbase + k7GroupsPerWarp + 2) + 4;
            const bool fast = !FROM_MAG && vec_ok && s0 + g_lo >= 0 && s0 + g_hi <= len;
            // rows travel global -> shared with cp.async (no registers held while in flight): the lane's
            // private ring keeps B200_SCAN7_RING rows ahead of the one being processed.  The ring is primed
            // BEFORE the group-closing row is fetched, so that the two DRAM latencies of a pass overlap.
            const char *gsrc = reinterpret_cast<const char *>(b32 + s0 + r0) + 48 * (k7Slots - 1);
            const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(ring) + 16u * (uint32_t)tid;
            constexpr uint32_t kRingStride = 16u * k7Threads;
            if (fast) {
#pragma unroll
                for (int d = 0; d < B200_SCAN7_RING; d++) {
#if !(B200_ABL & 1)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring0 + d * kRingStride),
                                 "l"(gsrc - 48 * d));
#endif
                    asm volatile("cp.async.commit_group;");
                }
                gsrc -= 48 * B200_SCAN7_RING;
            }
            {
                const int rb = k7Group * (G + 1);                  // first row of the next group
                Row rbnd;
                if (fast) {
                    const int4 v = __ldg(reinterpret_cast<const int4 *>(b32 + s0 + rb));
                    rbnd.q0 = mag_pair_fast2((uint32_t)v.x, (uint32_t)v.z);
                    rbnd.q1 = mag_pair_fast2((uint32_t)v.y, (uint32_t)v.w);
                } else {
                    rbnd = row_slow<FROM_MAG>(p, b32, d16, s0 + rb, i0 + rb, len, cs, b);
                }
                float x, y;
                f2_unpack(rbnd.q0, st.pm0, st.pm2);
                f2_unpack(rbnd.q1, st.pm1, y);
                (void)x;
            }
            uint16_t *mrow = mag + r0 + 12 * (k7Slots - 1);
            if (fast) {
                uint32_t rp = ring0;
#pragma unroll 1   // measured: 2 slots per iteration spill at 72 registers and cost 4 % (0.390 vs 0.373 ms)
                for (int k = k7Slots - 1; k >= 0; k--) {
                    asm volatile("cp.async.wait_group %0;" ::"n"(B200_SCAN7_RING - 1));
                    uint32_t x, y, z, ww;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(ww) : "r"(rp));
#if !(B200_ABL & 1)
                    if (k >= B200_SCAN7_RING)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(rp), "l"(gsrc));
#endif
                    asm volatile("cp.async.commit_group;");
                    gsrc -= 48;
                    rp += kRingStride;
                    if (rp == ring0 + B200_SCAN7_RING * kRingStride)
                        rp = ring0;
                    Row r;
                    r.q0 = mag_pair_abl(x, z);
                    r.q1 = mag_pair_abl(y, ww);
                    dense_slot(st, r, is_c0, src_lane, mrow, own);
                    mrow -= 12;
                }
            } else {
#pragma unroll 1
                for (int k = k7Slots - 1; k >= 0; k--) {
                    const int rr = r0 + 12 * k;
                    const Row r = row_slow<FROM_MAG>(p, b32, d16, s0 + rr, i0 + rr, len, cs, b);
                    dense_slot(st, r, is_c0, src_lane, mrow, own);
                    mrow -= 12;
                }
            }
            if (own) {
                // accumulator (e, f) is halfword G of plane row f*12 + 4c + e
#pragma unroll
                for (int f = 0; f < 7; f++)
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        planes16[2 * ((f * 12 + 4 * c + e) * WP) + G] = (uint16_t)st.acc[e][f];
            }
        }
    }
    __syncthreads();
#if B200_SCAN7_STOP == 1
    {   // measurement build: the dense phase alone (keep its stores observable)
        uint32_t x = 0;
        for (int i = tid; i < 7 * 12 * WP; i += k7Threads)
            x ^= planes[i];
        if (tid == 0)
            p.tile_dir[tile] = make_uint2(0u, 0u);
        if (x == 0x12345678u)
            p.rec[tid] = x + mag[tid];
        return;
    }
#endif

    // ---- P3a: preamble templates (demod_2400.rs:221-317).  Warp 0, lane = word column w, residue
    // static: the 32 positions 12*(32w+bit)+rho.  Edge bit at offset s of such a position is bit
    // (bit + carry) of row (rho+s) mod 12, carry = (rho+s) / 12.  Output per (rho, w): the match mask
    // and the template case as three bit planes.
    if (tid == 0)
        s_count = 0;
    if (warp == 0)
        p3a_rows<0, 12>(planes, WP, nwq, lane, masks, P.Wrow);
    __syncthreads();
    // ---- P3b: expand the match masks into one list (order is irrelevant: the gates only set
    // survivor bits): a warp takes 32 (rho, w) words per round
    {
        const uint4 *min4 = reinterpret_cast<const uint4 *>(masks);
        const int nwords = 12 * P.Wrow;
        for (int i0w = 32 * warp; i0w < nwords; i0w += 32 * k7Warps) {
            const int i = i0w + lane;
            const int rho = (int)(((uint32_t)i * (uint32_t)P.inv_Wrow) >> 16), w = i - rho * P.Wrow;   // i / Wrow (exact, checked for every Wrow <= 23)
            uint4 m = make_uint4(0u, 0u, 0u, 0u);
            if (i < nwords && w < nwq)
                m = min4[i];
            const int cnt = __popc(m.x);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o)
                    incl += t;
            }
            int base = 0;
            if (lane == 31 && incl)
                base = (int)atomicAdd(&s_count, (uint32_t)incl);
            base = __shfl_sync(0xffffffffu, base, 31);
            // (mask word, bit) entries; the gate thread decodes.  Entries past the capacity are dropped here:
            // the overflow pass below then gates every match in place.
            const int off = base + incl - cnt;
            const int take = min(cnt, max(P.list_cap - off, 0));
            uint16_t *lp = list + off;
            uint32_t any = m.x;
#pragma unroll 1
            for (int k = 0; k < take; k++) {
                const uint32_t bit = 31u - (uint32_t)__clz(any);
                any ^= 1u << bit;
                lp[k] = (uint16_t)((uint32_t)i | (bit << 10));
            }
        }
    }
    __syncthreads();
    if (s_count > (uint32_t)P.list_cap) {
        // more matches than the list holds (pathological input): gate every match in place.  Matches
        // that also made it into the list are evaluated twice, which is harmless (a gate only sets a bit).
        const uint4 *min4 = reinterpret_cast<const uint4 *>(masks);
        for (int i = tid; i < 12 * P.Wrow; i += k7Threads) {
            const int rho = (int)(((uint32_t)i * (uint32_t)P.inv_Wrow) >> 16), w = i - rho * P.Wrow;
            if (w >= nwq)
                continue;
            const uint4 m = min4[i];
            for (uint32_t any = m.x; any; any &= any - 1) {
                const int bit = __ffs(any) - 1;
                const uint32_t cs = ((m.y >> bit) & 1u) | (((m.z >> bit) & 1u) << 1) | (((m.w >> bit) & 1u) << 2);
                gate_eval7_cold(mag, surv, 12 * (32 * w + bit) + rho, cs, npos);
            }
        }
    }
    // ---- P3c: SNR and quiet-zone gates, one match per thread
    {
        const uint4 *min4 = reinterpret_cast<const uint4 *>(masks);
        const int n = min((int)s_count, P.list_cap);
        for (int g = tid; g < n; g += k7Threads) {
            const uint32_t e = list[g];
            const int i = (int)(e & 0x3ffu), bit = (int)(e >> 10);
            const uint4 m = min4[i];
            const uint32_t cs = ((m.y >> bit) & 1u) | (((m.z >> bit) & 1u) << 1) | (((m.w >> bit) & 1u) << 2);
            const int rho = (int)(((uint32_t)i * (uint32_t)P.inv_Wrow) >> 16), w = i - rho * P.Wrow;
            gate_eval_bf(mag, surv, 12 * (32 * w + bit) + rho, cs, npos);
        }
    }
    __syncthreads();

#if B200_SCAN7_STOP == 2
    if (tid == 0)
        p.tile_dir[tile] = make_uint2(0u, 0u);
    if (surv[tid] == 0x12345678u)
        p.rec[tid] = 1u;
    return;
#endif
    // ---- P4a: count survivors, reserve pool space (positions are emitted in ascending j).  Warp 0
    // alone: lane l owns the survivor words 8l..8l+7 (nw <= 256), one warp scan, no block barrier.
    uint32_t wv[8];
    int my_off = 0;
    if (warp == 0) {
        int cnt = 0;
#pragma unroll
        for (int h = 0; h < 8; h++) {
            const int wi = 8 * lane + h;
            wv[h] = (wi < P.nw) ? surv[wi] : 0u;
            cnt += __popc(wv[h]);
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        my_off = incl - cnt;
        {   // exclusive survivor offset of every pair of words: all warps fill the candidate list
            int o = my_off;
#pragma unroll
            for (int h2 = 0; h2 < 4; h2++) {
                s_pair_off[4 * lane + h2] = (uint16_t)o;
                o += __popc(wv[2 * h2]) + __popc(wv[2 * h2 + 1]);
            }
        }
        if (lane == 31) {
            const uint32_t total = (uint32_t)incl;
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

    // ---- P4b: five try-phases per survivor, in windows of k7CandCap survivors (as v6)
    const int C = (int)s_count;
    CrcLanes crcl;      // CRC-24 field tables, one entry per lane (looked up by warp shuffle)
#pragma unroll
    for (int c = 0; c < kLaneTabs; c++)
        crcl.t[c] = __ldg(p.crc_lanes + 32 * c + lane);
    const unsigned long long ord_buf = (p.ord_first + (unsigned long long)b * p.ord_stride) << 20;
    for (int win = 0; win < C; win += k7CandCap) {
        {
            int off = s_pair_off[tid];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int wi = 2 * tid + h;
                uint32_t wv2 = (wi < P.nw) ? surv[wi] : 0u;
                if (C <= k7CandCap) {            // the usual case, one window: no range test per survivor
                    while (wv2) {
                        const int bit = __ffs(wv2) - 1;
                        wv2 &= wv2 - 1;
                        cand[off++] = (uint16_t)(wi * 32 + bit);
                    }
                } else {
                    while (wv2) {
                        const int bit = __ffs(wv2) - 1;
                        wv2 &= wv2 - 1;
                        if (off >= win && off < win + k7CandCap)
                            cand[off - win] = (uint16_t)(wi * 32 + bit);
                        off++;
                    }
                }
            }
        }
        __syncthreads();
        const int Cw = min(k7CandCap, C - win);
        uint32_t *rec_w = p.rec + 6ull * (s_base + (uint32_t)win);
        for (int item = tid; item < 5 * Cw; item += k7Threads) {
            const int ci = item / 5, tt = item - 5 * ci;
            const int jl = cand[ci];
            // demod_2400.rs:158-160: P0 = 5*(mi+19) + try_phase, try_phase = 4+tt
            const uint32_t A = (uint32_t)(jl + kHaloFront + 19);
            const uint32_t qA = A / 12u, rA = A - 12u * qA;
            const uint32_t *lrow = lut + (25u * rA + 5u * (uint32_t)tt);      // (unsigned: one IMAD.WIDE, no sign extension)
            uint32_t f[5];
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const uint32_t e = __ldg(lrow + r);
                const uint32_t q = qA + (e >> 16);
                const uint32_t *stp = planes + (e & 0xffffu) + (q >> 5);
                // (the funnel shift takes its amount modulo 32)
                f[r] = __funnelshift_r(stp[0], stp[1], q) & (r < 2 ? 0x7fffffu : 0x3fffffu);
            }
            if (tt == 0)
                rec_w[6 * ci] = (uint32_t)(tile_start + jl);
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
            if (cls) {
                const uint32_t slot = cls == 1 ? basel + (uint32_t)__popc(ml & lt)
                                               : (uint32_t)(k7FieldItems - 1) - (bases + (uint32_t)__popc(ms & lt));
                uint32_t *o = fb + 5 * slot;
                o[0] = f[0];
                o[1] = f[1];
                o[2] = f[2] | (((uint32_t)item & 0x3ffu) << 22);
                o[3] = f[3] | (((uint32_t)item >> 10) << 22);
                o[4] = f[4];
            } else {
                rec_w[6 * ci + 1 + tt] = wd;
            }
        }
        __syncthreads();
        {
            // long items, then short ones: each loop has a warp-uniform trip count, so the table shuffles
            // inside are executed by whole warps (lanes past the end carry zeros)
            const int nl = (int)s_nlong, ns = (int)s_nshort;
            for (int g0 = 32 * warp; g0 < nl; g0 += k7Threads) {
                const int g = g0 + lane;
                const bool valid = g < nl;
                const uint32_t *o = fb + 5 * (valid ? g : 0);
                uint32_t f[5] = {o[0], o[1], o[2], o[3], o[4]};
                const int item = (int)((f[2] >> 22) | ((f[3] >> 22) << 10));
                f[2] &= 0x3fffffu;
                f[3] &= 0x3fffffu;
                const uint32_t syn = syn112_fields_sh(crcl, f);
                if (valid) {
                    const uint32_t df = df_of_fields(f);
                    uint32_t wd;
                    if (df == 17 || df == 18)              // mode_s/mod.rs:91-109
                        wd = syn ? 0u : (((df == 17 ? K_DF17 : K_DF18) << 29) | msg_bits<8, 24>(f));
                    else                                    // :110-134
                        wd = (K_PAR_LONG << 29) | syn;
                    const int ci = item / 5, tt = item - 5 * ci;
                    rec_w[6 * ci + 1 + tt] = wd;
                    const uint32_t kind = wd >> 29;
                    if (kind == K_DF17 || kind == K_DF18) {
                        const uint32_t key = (wd & 0xffffffu) | (kind == K_DF18 ? B200ADSB_ICAO_FILTER_ADSB_NT : 0u);
                        const uint32_t j = (uint32_t)(tile_start + cand[ci]);
                        event_add(p.ev_keys, p.ev_ord, p.ev_used, p.ev_mask, p.counters, key,
                                  ord_buf | ((unsigned long long)j << 3) | (unsigned long long)tt);
                    }
                }
            }
            for (int g0 = 32 * warp; g0 < ns; g0 += k7Threads) {
                const int g = g0 + lane;
                const bool valid = g < ns;
                const uint32_t *o = fb + 5 * (uint32_t)(k7FieldItems - 1 - (valid ? g : 0));
                uint32_t f[5] = {o[0], o[1], o[2], o[3], o[4]};
                const int item = (int)((f[2] >> 22) | ((f[3] >> 22) << 10));
                f[2] &= 0x3fffffu;
                f[3] &= 0x3fffffu;
                const uint32_t syn = syn56_fields_sh(crcl, f);
                if (valid) {
                    const uint32_t df = df_of_fields(f);
                    uint32_t wd;
                    if (df == 11)                           // :73-90
                        wd = (syn & 0xffff80u) ? 0u
                                               : ((((syn & 0x7f) ? K_DF11_IID : K_DF11_IID0) << 29) | msg_bits<8, 24>(f));
                    else                                    // :56-72
                        wd = (K_PAR_SHORT << 29) | syn;
                    const int ci = item / 5, tt = item - 5 * ci;
                    rec_w[6 * ci + 1 + tt] = wd;
                    if ((wd >> 29) == K_DF11_IID0) {
                        const uint32_t j = (uint32_t)(tile_start + cand[ci]);
                        event_add(p.ev_keys, p.ev_ord, p.ev_used, p.ev_mask, p.counters, wd & 0xffffffu,
                                  ord_buf | ((unsigned long long)j << 3) | (unsigned long long)tt);
                    }
                }
            }
        }
        if (win + k7CandCap >= C)
            break;                       // last window: nothing left to synchronise with
        __syncthreads();
        if (tid == 0) {
            s_nlong = 0;
            s_nshort = 0;
        }
    }
}

}  // namespace b200
