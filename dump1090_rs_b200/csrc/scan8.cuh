// dump1090_rs_b200/csrc/scan8.cuh -- stage 1 as two kernels (sm_100a):
//
//   dense8_kernel   the streaming part.  IQ -> u16 magnitude, rising/falling edge bits and the
//                   five PPM correlator signs of EVERY sample, written as mod-12 de-interleaved
//                   bit planes (src/utils.rs:43-58, src/demod_2400.rs:62-83,221-317).  Warps are
//                   independent (no block barrier, no shared memory besides the cp.async row ring),
//                   there is no tile halo: the kernel is a pure map over the sample stream.
//                   4 B/sample read, 2 B (magnitude) + 7/8 B (planes) written.
//   sparse8_kernel  everything that touches a few percent of the positions: preamble templates
//                   on the edge planes, SNR / quiet-zone gates, five try-phases per survivor,
//                   CRC-24, records + ICAO add-events (src/demod_2400.rs:127-189,
//                   src/mode_s/mod.rs:34-139).  Reads the planes / magnitudes through L1/L2.
//
// Global index i' = data index + 2 (so that IQ rows are 16-byte aligned: sample = i' - 328).
// plane[f][rho] bit q  <->  i' = 12 q + rho;  position j (demod_2400.rs:121) <-> i' = j + 2.
#pragma once
#include "scan7.cuh"

namespace b200 {

constexpr int k8MaxCols = 32;                       // word columns (384 positions each) per sparse tile
constexpr int k8ListCap = 1024;
#ifndef B200_DENSE8_MIN_BLOCKS
#define B200_DENSE8_MIN_BLOCKS 6
#endif
#ifndef B200_SPARSE8_MIN_BLOCKS
#define B200_SPARSE8_MIN_BLOCKS 8
#endif

struct Scan8Geom {
    int NGb, WQ, MS, nblk;
    __host__ __device__ explicit Scan8Geom(int spb)
    {
        NGb = (kTrailing + kHaloFront + spb + kHaloTot + k7Group - 1) / k7Group + 1;
        WQ = (NGb + 1) / 2 + 2;                     // words per plane row (+ zero pad read by funnel shifts)
        MS = NGb * k7Group + 32;                    // u16 magnitudes per buffer (+ pad)
        nblk = (NGb + k7GroupsPerPass - 1) / k7GroupsPerPass;
    }
};

struct Scan8Params {
    ScanParams s;
    uint16_t *magG;        // [chunk][MS]
    uint32_t *planesG;     // [chunk][7 * 12][WQ]
    uint32_t b_off;        // first buffer of this chunk (index into s.in / s.lengths / s.tile_dir)
    int NGb, WQ, MS, nblk;
    int nW;                // word columns per sparse tile (s.T = 384 * nW)
};

// ================================================================== dense
template <bool FROM_MAG>
__global__ void __launch_bounds__(k7Threads, B200_DENSE8_MIN_BLOCKS) dense8_kernel(const Scan8Params P)
{
    __shared__ __align__(16) unsigned char ring[(B200_SCAN7_RING > 0 ? B200_SCAN7_RING : 1) * k7Threads * 16];
    const ScanParams &p = P.s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t bl = blockIdx.x / (uint32_t)P.nblk;              // buffer inside the chunk
    const int blk = (int)(blockIdx.x - bl * (uint32_t)P.nblk);
    const uint32_t b = P.b_off + bl;
    const int len = p.lengths ? (int)min(p.lengths[b], p.spb) : (int)p.spb;
    const int NG = P.NGb;
    int prev_len = 0;
    const uint32_t *prev = FROM_MAG ? nullptr
                                    : carry_source(p.in, p.stride, p.lengths, p.spb, b, p.carry, p.tail, &prev_len);
    const uint32_t *b32 = reinterpret_cast<const uint32_t *>(p.in) + (unsigned long long)b * p.stride;
    const uint16_t *d16 = reinterpret_cast<const uint16_t *>(p.in) + (unsigned long long)b * p.stride;
    const int s0 = -(kTrailing + kHaloFront);     // sample index of i' = 0
    const int i0 = -kHaloFront;                    // data index of i' = 0
    uint16_t *magb = P.magG + (size_t)bl * P.MS;
    uint16_t *planes16 = reinterpret_cast<uint16_t *>(P.planesG + (size_t)bl * 84 * P.WQ);
    const int gl = lane / 3, c = lane - 3 * gl;
    const int src_lane = (c < 2) ? min(lane + 1, 31) : lane - 2;
    const bool is_c0 = c == 0;
    const int gbase = blk * k7GroupsPerPass + warp * k7GroupsPerWarp;
    if (gbase >= NG)
        return;
    const int G = gbase + gl;
    const bool own = lane < 30 && G < NG;
    const int r0 = k7Group * G + 4 * c;            // i' of slot 0
    DenseState st;
#pragma unroll
    for (int e = 0; e < 4; e++)
#pragma unroll
        for (int f = 0; f < 7; f++)
            st.acc[e][f] = 0;
    const int g_lo = k7Group * gbase, g_hi = k7Group * (gbase + k7GroupsPerWarp + 2) + 4;
    const bool fast = !FROM_MAG && p.vec_ok && s0 + g_lo >= 0 && s0 + g_hi <= len;
    {
        const int rb = k7Group * (G + 1);
        Row rbnd;
        if (fast) {
            const int4 v = __ldg(reinterpret_cast<const int4 *>(b32 + s0 + rb));
            rbnd.q0 = mag_pair_fast2((uint32_t)v.x, (uint32_t)v.z);
            rbnd.q1 = mag_pair_fast2((uint32_t)v.y, (uint32_t)v.w);
        } else {
            rbnd = row_slow<FROM_MAG>(p, b32, d16, s0 + rb, i0 + rb, len, prev, prev_len);
        }
        float y;
        f2_unpack(rbnd.q0, st.pm0, st.pm2);
        f2_unpack(rbnd.q1, st.pm1, y);
    }
    uint16_t *mrow = magb + r0 + 12 * (k7Slots - 1);
    if (fast) {
#if B200_SCAN7_RING > 0
        const char *gsrc = reinterpret_cast<const char *>(b32 + s0 + r0) + 48 * (k7Slots - 1);
        const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(ring) + 16u * (uint32_t)tid;
        constexpr uint32_t kRingStride = 16u * k7Threads;
#pragma unroll
        for (int d = 0; d < B200_SCAN7_RING; d++) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring0 + d * kRingStride), "l"(gsrc - 48 * d));
            asm volatile("cp.async.commit_group;");
        }
        gsrc -= 48 * B200_SCAN7_RING;
        uint32_t rp = ring0;
#pragma unroll 2
        for (int k = k7Slots - 1; k >= 0; k--) {
            asm volatile("cp.async.wait_group %0;" ::"n"(B200_SCAN7_RING - 1));
            uint32_t x, y, z, ww;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(ww) : "r"(rp));
            if (k >= B200_SCAN7_RING)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(rp), "l"(gsrc));
            asm volatile("cp.async.commit_group;");
            gsrc -= 48;
            rp += kRingStride;
            if (rp == ring0 + B200_SCAN7_RING * kRingStride)
                rp = ring0;
            Row r;
            r.q0 = mag_pair_fast2(x, z);
            r.q1 = mag_pair_fast2(y, ww);
            dense_slot(st, r, is_c0, src_lane, mrow, own);
            mrow -= 12;
        }
#else
        const int4 *src = reinterpret_cast<const int4 *>(b32 + s0 + r0) + 3 * (k7Slots - 1);
        int4 va = __ldg(src), vb;
#pragma unroll 1
        for (int k = k7Slots / 2 - 1; k >= 0; k--) {
            vb = __ldg(src - 3);
            Row r;
            r.q0 = mag_pair_fast2((uint32_t)va.x, (uint32_t)va.z);
            r.q1 = mag_pair_fast2((uint32_t)va.y, (uint32_t)va.w);
            dense_slot(st, r, is_c0, src_lane, mrow, own);
            src -= 6;
            if (k > 0)
                va = __ldg(src);
            r.q0 = mag_pair_fast2((uint32_t)vb.x, (uint32_t)vb.z);
            r.q1 = mag_pair_fast2((uint32_t)vb.y, (uint32_t)vb.w);
            dense_slot(st, r, is_c0, src_lane, mrow - 12, own);
            mrow -= 24;
        }
#endif
    } else {
#pragma unroll 1
        for (int k = k7Slots - 1; k >= 0; k--) {
            const int rr = r0 + 12 * k;
            const Row r = row_slow<FROM_MAG>(p, b32, d16, s0 + rr, i0 + rr, len, prev, prev_len);
            dense_slot(st, r, is_c0, src_lane, mrow, own);
            mrow -= 12;
        }
    }
    if (own) {
        // accumulator (e, f) is halfword G of plane row f*12 + 4c + e
        uint16_t *dst = planes16 + 2 * (size_t)((4 * c) * P.WQ) + G;
        const int rs = 2 * P.WQ;
#pragma unroll
        for (int f = 0; f < 7; f++)
#pragma unroll
            for (int e = 0; e < 4; e++)
                dst[(f * 12 + e) * rs] = (uint16_t)st.acc[e][f];
    }
}

// ================================================================== sparse
struct Sparse8Smem {
    size_t off_masks, off_list, off_surv, off_cand, off_fb, bytes;
    __host__ __device__ explicit Sparse8Smem(int nW)
    {
        size_t o = 0;
        off_masks = o;
        o += (size_t)12 * nW * 16;
        off_list = o;
        o += (size_t)k8ListCap * 4;
        off_surv = o;
        o += (size_t)12 * nW * 4;
        off_cand = o;
        o += (size_t)k7CandCap * 2;
        o = (o + 15) & ~(size_t)15;
        off_fb = o;
        o += (size_t)k7FieldItems * 5 * 4;
        bytes = o;
    }
};

__device__ __forceinline__ void gate8(const uint16_t *magt, uint32_t *surv, int il, uint32_t cs, int j0, int len)
{
    // il: i' local to the tile; position j = j0 + il
    const int j = j0 + il;
    if (j < 0 || j >= len)
        return;
    const uint16_t *pp = magt + il;
    const int a = __ldg(pp + 1), h = __ldg(pp + 2), b = __ldg(pp + 3), e = __ldg(pp + 4), n5 = __ldg(pp + 5),
              n6 = __ldg(pp + 6), n7 = __ldg(pp + 7), n8 = __ldg(pp + 8);
    const int c = __ldg(pp + 9), f = __ldg(pp + 10), g = __ldg(pp + 11), d = __ldg(pp + 12);
    const int bc = b + c, ef = e + f;
    const int H = a + d + (cs < 3 ? bc : 0) + (cs >= 2 ? ef : 0) + (cs == 0 ? g : 0) + (cs == 4 ? h : 0);
    const int S = (cs < 4 ? a : 0) + (cs < 2 ? bc : 0) + (cs >= 1 ? d : 0) + (cs >= 3 ? ef : 0);
    const int N = n6 + n7 + (((0x0Bu >> cs) & 1u) ? n5 : 0) + (((0x1Au >> cs) & 1u) ? n8 : 0);
    if (2 * S < 3 * N)            // demod_2400.rs:129
        return;
    const int mx = max(max(max(n5, n6), max(n7, n8)),
                       max(max(max((int)__ldg(pp + 14), (int)__ldg(pp + 15)), max((int)__ldg(pp + 16), (int)__ldg(pp + 17))),
                           (int)__ldg(pp + 18)));
    if (mx >= (H >> 2))           // demod_2400.rs:135-146
        return;
    atomicOr(&surv[il >> 5], 1u << (il & 31));
}
__device__ __noinline__ void gate8_cold(const uint16_t *magt, uint32_t *surv, int il, uint32_t cs, int j0, int len)
{
    gate8(magt, surv, il, cs, j0, len);
}

__global__ void __launch_bounds__(k7Threads, B200_SPARSE8_MIN_BLOCKS) sparse8_kernel(const Scan8Params P)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const ScanParams &p = P.s;
    const Sparse8Smem L(P.nW);
    uint32_t *masks = reinterpret_cast<uint32_t *>(smem + L.off_masks);   // [12][nWt] x (match, case planes)
    uint32_t *list = reinterpret_cast<uint32_t *>(smem + L.off_list);
    uint32_t *surv = reinterpret_cast<uint32_t *>(smem + L.off_surv);     // bit il <-> i' local
    uint16_t *cand = reinterpret_cast<uint16_t *>(smem + L.off_cand);
    uint32_t *fb = reinterpret_cast<uint32_t *>(smem + L.off_fb);
    const uint32_t *tabs = p.crc_tabs;
    const uint32_t *lut = p.lut;
    __shared__ uint32_t s_base, s_count, s_ok, s_nlong, s_nshort;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tl = blockIdx.x;                                   // tile inside the chunk
    const uint32_t bl = tl / (uint32_t)p.tiles_per_buffer;
    const int kt = (int)(tl - bl * (uint32_t)p.tiles_per_buffer);
    const uint32_t b = P.b_off + bl;
    const uint32_t tile = b * (uint32_t)p.tiles_per_buffer + (uint32_t)kt;
    const int len = p.lengths ? (int)min(p.lengths[b], p.spb) : (int)p.spb;
    const int c0 = P.nW * kt;                                          // first word column
    const int ip0 = 384 * c0;                                          // i' of the tile's first position
    const int j0 = ip0 - kHaloFront;                                   // position of il = 0
    if (j0 >= len) {
        if (tid == 0)
            p.tile_dir[tile] = make_uint2(0u, 0u);
        return;
    }
    const int nWt = min(P.nW, (len + kHaloFront - ip0 + 383) / 384);   // columns that hold positions
    const int WQ = P.WQ;
    const uint32_t *planes = P.planesG + (size_t)bl * 84 * WQ;
    const uint16_t *magt = P.magG + (size_t)bl * P.MS + ip0;
    const int nws = 12 * nWt;                                          // survivor words in use

    for (int i = tid; i < nws; i += k7Threads)
        surv[i] = 0;
    // ---- P3a: preamble templates (demod_2400.rs:221-317).  Warp 0, lane = word column, residue
    // static; edge bit at offset s of position (rho, bit) is bit (bit + carry) of row (rho+s) mod 12.
    if (warp == 0) {
        const uint32_t *Rp = planes + 5 * 12 * WQ + c0, *Fp = planes + 6 * 12 * WQ + c0;
        if (tid == 0)
            s_count = 0;
        const int w = lane;
        if (w < nWt) {
            uint32_t XR[24], XF[24];
#pragma unroll
            for (int t = 0; t < 12; t++) {
                const uint32_t r0 = __ldg(Rp + t * WQ + w), r1 = __ldg(Rp + t * WQ + w + 1);
                const uint32_t f0 = __ldg(Fp + t * WQ + w), f1 = __ldg(Fp + t * WQ + w + 1);
                XR[t] = r0;
                XF[t] = f0;
                XR[t + 12] = __funnelshift_r(r0, r1, 1);
                XF[t + 12] = __funnelshift_r(f0, f1, 1);
            }
            uint4 *mout = reinterpret_cast<uint4 *>(masks) + w;
#pragma unroll
            for (int rho = 0; rho < 12; rho++) {
#define ER(s) XR[rho + (s)]
#define EF(s) XF[rho + (s)]
                const uint32_t quick = ER(0) & EF(12);   // p0 < p1 && p12 > p13 (:221)
                const uint32_t T3 = EF(1) & ER(2) & EF(3) & ER(8) & EF(9) & ER(10);
                const uint32_t T4 = EF(1) & ER(2) & EF(3) & ER(8) & EF(9) & ER(11);
                const uint32_t T5 = EF(1) & ER(2) & EF(4) & ER(8) & EF(10) & ER(11);
                const uint32_t T6 = EF(1) & ER(3) & EF(4) & ER(9) & EF(10) & ER(11);
                const uint32_t T7 = EF(2) & ER(3) & EF(4) & ER(9) & EF(10) & ER(11);
#undef ER
#undef EF
                const uint32_t c1 = T4 & ~T3, c2 = T5 & ~(T3 | T4), c3 = T6 & ~(T3 | T4 | T5);
                mout[rho * nWt] = make_uint4(quick & (T3 | T4 | T5 | T6 | T7), c1 | c3, c2 | c3, ~(T3 | T4 | T5 | T6));
            }
        }
    }
    __syncthreads();
    // ---- P3b: expand the match masks into one list (order irrelevant), 32 (rho, w) words per warp round
    {
        const uint4 *min4 = reinterpret_cast<const uint4 *>(masks);
        for (int i0w = 32 * warp; i0w < nws; i0w += 32 * k7Warps) {
            const int i = i0w + lane;
            uint4 m = make_uint4(0u, 0u, 0u, 0u);
            if (i < nws)
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
            int off = base + incl - cnt;
            const int rho = i / nWt, w = i - rho * nWt;
            const int il0 = 384 * w + rho;
            uint32_t any = m.x;
            while (any) {
                const int bit = __ffs(any) - 1;
                any &= any - 1;
                const uint32_t cs = ((m.y >> bit) & 1u) | (((m.z >> bit) & 1u) << 1) | (((m.w >> bit) & 1u) << 2);
                const int il = il0 + 12 * bit;
                if (off < k8ListCap)
                    list[off] = (uint32_t)il | (cs << 16);
                else
                    gate8_cold(magt, surv, il, cs, j0, len);
                off++;
            }
        }
    }
    __syncthreads();
    // ---- P3c: SNR and quiet-zone gates, one match per thread
    {
        const int n = min((int)s_count, k8ListCap);
        for (int g = tid; g < n; g += k7Threads) {
            const uint32_t e = list[g];
            gate8(magt, surv, (int)(e & 0xffffu), e >> 16, j0, len);
        }
    }
    __syncthreads();

    // ---- P4a: count survivors, reserve pool space (ascending j).  Warp 0: lane l owns the
    // survivor words 12l..12l+11 (nws <= 384)
    uint32_t wv[12];
    int my_off = 0;
    if (warp == 0) {
        int cnt = 0;
#pragma unroll
        for (int h = 0; h < 12; h++) {
            const int wi = 12 * lane + h;
            wv[h] = (wi < nws) ? surv[wi] : 0u;
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

    // ---- P4b: five try-phases per survivor, in windows of k7CandCap survivors
    const int C = (int)s_count;
    const unsigned long long ord_buf = (p.ord_first + (unsigned long long)b * p.ord_stride) << 20;
    for (int win = 0; win < C; win += k7CandCap) {
        if (warp == 0) {
            int off = my_off;
#pragma unroll
            for (int h = 0; h < 12; h++) {
                uint32_t wv2 = wv[h];
                while (wv2) {
                    const int bit = __ffs(wv2) - 1;
                    wv2 &= wv2 - 1;
                    if (off >= win && off < win + k7CandCap)
                        cand[off - win] = (uint16_t)((12 * lane + h) * 32 + bit);
                    off++;
                }
            }
        }
        __syncthreads();
        const int Cw = min(k7CandCap, C - win);
        uint32_t *rec_w = p.rec + 6ull * (s_base + (uint32_t)win);
        for (int item = tid; item < 5 * Cw; item += k7Threads) {
            const int ci = item / 5, tt = item - 5 * ci;
            const int il = cand[ci];
            // demod_2400.rs:158-160: P0 = 5*(i'+19) + try_phase, try_phase = 4+tt
            const int A = ip0 + il + 19;
            const int qA = A / 12, rA = A - 12 * qA;
            const uint32_t *lrow = lut + 25 * rA + 5 * tt;
            uint32_t f[5];
#pragma unroll
            for (int r = 0; r < 5; r++) {
                const uint32_t e = __ldg(lrow + r);
                const int q = qA + (int)(e >> 16);
                const uint32_t *stp = planes + (e & 0xffffu) + (q >> 5);
                f[r] = __funnelshift_r(__ldg(stp), __ldg(stp + 1), q & 31) & (r < 2 ? 0x7fffffu : 0x3fffffu);
            }
            if (tt == 0)
                rec_w[6 * ci] = (uint32_t)(j0 + il);
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
            const int nl = (int)s_nlong, ns = (int)s_nshort;
            for (int g = tid; g < nl + ns; g += k7Threads) {
                const bool is_long = g < nl;
                const uint32_t slot = is_long ? (uint32_t)g : (uint32_t)(k7FieldItems - 1 - (g - nl));
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
                const int ci = item / 5, tt = item - 5 * ci;
                rec_w[6 * ci + 1 + tt] = wd;
                const uint32_t kind = wd >> 29;
                if (kind == K_DF11_IID0 || kind == K_DF17 || kind == K_DF18) {
                    const uint32_t key = (wd & 0xffffffu) | (kind == K_DF18 ? B200ADSB_ICAO_FILTER_ADSB_NT : 0u);
                    const uint32_t j = (uint32_t)(j0 + cand[ci]);
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

}  // namespace b200
