// dump1090_rs_b200/csrc/b200adsb.cu -- C ABI (include/b200adsb.h) over the kernels.
//
// Host side of libb200adsb.so: context (device buffers, stream, ICAO filter state
// resident on the GPU), the staged pipeline scan -> [event exchange] -> finalise ->
// resolve -> ordered emit -> commit, and the host<->device plumbing.  No torch, no
// CPU compute path: if CUDA is unavailable every entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "kernels.cuh"
#include "scan7.cuh"

using namespace b200;

namespace {

constexpr uint32_t kEvSlots = 1u << 16;
constexpr int kRedo = 1;   // internal status: the optimistic scan overflowed, redo with a larger pool
constexpr size_t kMaxStageBytes = 2ull << 30;   // host batch API: IQ staged per pass

struct EventPair {
    cudaEvent_t a, b;
};

struct Pending {
    bool active = false;
    const void *in = nullptr;
    bool from_mag = false;
    const uint32_t *lengths = nullptr;
    uint32_t n_buffers = 0, spb = 0, n_tiles = 0;
    unsigned long long stride = 0;
    int T = 0, tpb = 0;
    unsigned long long ord_first = 0, ord_stride = 1;
    const uint8_t *msgs = nullptr;   // message-level API
};

}  // namespace

struct b200adsb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    bool own_stream = false;
    int tile_opt = 0, pool_shift = 5, profile = 0, h2d_chunk = 64, carry = 0;
    uint32_t *d_tails = nullptr;   // carry mode: last 326 IQ samples of the stream, double buffered
                                   // (2 x kTailWords; counters[C_TAILCUR] selects the current one)

    uint32_t *d_counters = nullptr, *h_counters = nullptr;
    uint32_t *d_members = nullptr;
    uint32_t *d_ev_keys = nullptr, *d_ev_used = nullptr, *d_ev_tmp = nullptr, *d_new_keys = nullptr;
    unsigned long long *d_ev_ord = nullptr;
    uint32_t *d_crc256 = nullptr, *d_lut = nullptr, *d_crc_lanes = nullptr;
    uint32_t h_lut[kLutWords];
    int lut_T = 0, lut_WP = 0;
    bool counters_clean = false;     // C_POOL/C_FLAGS/C_CAND already zero (cleared by the last commit kernel)
    uint32_t *d_scalar = nullptr;

    uint32_t *d_rec = nullptr, *d_emit_info = nullptr;
    int32_t *d_rec_score = nullptr;
    size_t pool_cap = 0;
    uint2 *d_tile_dir = nullptr;
    uint32_t *d_tile_emit = nullptr, *d_cta_sum = nullptr, *d_bloom = nullptr;
    size_t tiles_cap = 0;

    void *d_stage = nullptr;
    size_t stage_bytes = 0;
    uint8_t *d_stage8 = nullptr;      // CU8 ingest: the 8-bit samples as they crossed PCIe
    size_t stage8_bytes = 0;
    b200adsb_frame *d_frames = nullptr;
    size_t frames_cap = 0;
    uint32_t *d_counts = nullptr;
    size_t counts_cap = 0;
    uint32_t *d_lengths = nullptr;
    size_t lengths_cap = 0;

    // receive-loop slots (b200adsb_demod_iq_batch_submit / _wait): device staging, frames, outcome, completion
    struct Slot {
        void *d_iq = nullptr;
        size_t iq_bytes = 0;
        uint32_t *d_lengths = nullptr;
        size_t lengths_cap = 0;
        b200adsb_frame *d_frames = nullptr;
        size_t frames_cap = 0;
        uint32_t *d_result = nullptr;
        cudaEvent_t done = nullptr;
        bool in_flight = false;
    } slots[2];

    unsigned long long next_ordinal = 0;   // stream position (buffers) for the fused entry points
    Pending cur;
    std::vector<EventPair> scan_events, other_events;
    std::vector<cudaEvent_t> chunk_events;
    b200adsb_timing timing{};
    char err[256] = {0};
};

namespace {

#define CK(ctx, call)                                                                     \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, \
                     cudaGetErrorString(e_));                                             \
            return B200ADSB_ERR_CUDA;                                                     \
        }                                                                                 \
    } while (0)

uint32_t crc_pow(int e)   // x^e mod G, G = x^24 + 0xFFF409
{
    uint32_t s = 1;
    for (int i = 0; i < e; i++) {
        s <<= 1;
        if (s & 0x1000000u)
            s ^= 0x1FFF409u;
    }
    return s;
}

// Field tables of kernels.cuh::a112/a56: field bit m of field r is message bit 5m+r.
//   a112(f) = sum_{m<=21} f[m] x^(107-5m),  a56(f) = sum_{m<=10} f[m] x^(51-5m)
void build_crc_tabs(uint32_t *t, uint32_t *t256, uint32_t *lanes = nullptr)
{
    auto fill = [&](uint32_t *dst, int entries, int m0, int e_base) {
        for (int v = 0; v < entries; v++) {
            uint32_t s = 0;
            for (int bit = 0; (1 << bit) < entries; bit++)
                if (v & (1 << bit))
                    s ^= crc_pow(e_base - 5 * (m0 + bit));
            dst[v] = s;
        }
    };
    fill(t, 256, 0, 107);
    fill(t + 256, 256, 8, 107);
    fill(t + 512, 64, 16, 107);
    fill(t + kTab56, 256, 0, 51);
    fill(t + kTab56 + 256, 8, 8, 51);
    // the same sums as shuffle tables (kernels.cuh::a112_sh / a56_sh): 5-bit chunks of a field
    if (lanes)
        for (int v = 0; v < 32; v++)
            for (int c = 0; c < kLaneTabs; c++) {
                uint32_t sx = 0;
                for (int bit = 0; bit < 5; bit++) {
                    const int m = 5 * (c < 5 ? c : c - 5) + bit;
                    if (((v >> bit) & 1) && (c < 5 ? m <= 21 : true))
                        sx ^= crc_pow((c < 5 ? 107 : 51) - 5 * m);
                }
                lanes[32 * c + v] = sx;
            }
    for (uint32_t i = 0; i < 256; i++) {   // src/crc.rs:3-260, generated from the polynomial
        uint32_t c = i << 16;
        for (int k = 0; k < 8; k++)
            c = (c & 0x800000u) ? ((c << 1) ^ 0xFFF409u) : (c << 1);
        t256[i] = c & 0xFFFFFFu;
    }
}

int bind(b200adsb_ctx *c)
{
    CK(c, cudaSetDevice(c->device));
    return B200ADSB_OK;
}

template <typename T>
int grow(b200adsb_ctx *c, T **p, size_t *cap, size_t need, size_t elem = sizeof(T))
{
    if (need <= *cap && *p)
        return B200ADSB_OK;
    if (*p)
        CK(c, cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    const size_t n = std::max<size_t>(need, 1);
    cudaError_t e = cudaMalloc((void **)p, n * elem);
    if (e != cudaSuccess) {
        snprintf(c->err, sizeof(c->err), "cudaMalloc(%zu bytes): %s", n * elem, cudaGetErrorString(e));
        return B200ADSB_ERR_NOMEM;
    }
    *cap = n;
    return B200ADSB_OK;
}

int ensure_pool(b200adsb_ctx *c, size_t need)
{
    if (need <= c->pool_cap && c->d_rec)
        return B200ADSB_OK;
    if (c->d_rec) CK(c, cudaFree(c->d_rec));
    if (c->d_emit_info) CK(c, cudaFree(c->d_emit_info));
    if (c->d_rec_score) CK(c, cudaFree(c->d_rec_score));
    c->d_rec = c->d_emit_info = nullptr;
    c->d_rec_score = nullptr;
    c->pool_cap = 0;
    cudaError_t e = cudaMalloc((void **)&c->d_rec, need * 24);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_emit_info, need * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_rec_score, need * 4);
    if (e != cudaSuccess) {
        snprintf(c->err, sizeof(c->err), "candidate pool (%zu records): %s", need, cudaGetErrorString(e));
        return B200ADSB_ERR_NOMEM;
    }
    c->pool_cap = need;
    return B200ADSB_OK;
}

int ensure_tiles(b200adsb_ctx *c, size_t n_tiles)
{
    if (n_tiles <= c->tiles_cap && c->d_tile_dir)
        return B200ADSB_OK;
    if (c->d_tile_dir) CK(c, cudaFree(c->d_tile_dir));
    if (c->d_tile_emit) CK(c, cudaFree(c->d_tile_emit));
    if (c->d_cta_sum) CK(c, cudaFree(c->d_cta_sum));
    c->d_cta_sum = nullptr;
    c->d_tile_dir = nullptr;
    c->d_tile_emit = nullptr;
    c->tiles_cap = 0;
    cudaError_t e = cudaMalloc((void **)&c->d_tile_dir, n_tiles * sizeof(uint2));
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_tile_emit, (n_tiles + 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_cta_sum, (n_tiles / 32 + 2) * 4);
    if (e != cudaSuccess) {
        snprintf(c->err, sizeof(c->err), "tile directory (%zu): %s", n_tiles, cudaGetErrorString(e));
        return B200ADSB_ERR_NOMEM;
    }
    c->tiles_cap = n_tiles;
    return B200ADSB_OK;
}

int pick_tile(const b200adsb_ctx *c, size_t n_buffers, size_t spb)
{
    if (c->tile_opt)
        return c->tile_opt;
    // tile sizes whose halo-extended length is a whole number of warp passes (10 groups of 192
    // samples); the largest that still gives every SM a few tiles
    static const int kTiles[] = {kDefaultTile, 3544, 1624};   // whole warps of 10 groups x 192 samples
    for (int ti = 0; ti < 3; ti++) {
        const int t = kTiles[ti];
        const size_t tiles = n_buffers * ((spb + t - 1) / t);
        if (tiles >= 592)
            return t;
    }
    return kTiles[2];
}

void prof_begin(b200adsb_ctx *c, std::vector<EventPair> &v)
{
    if (!c->profile)
        return;
    EventPair p;
    cudaEventCreate(&p.a);
    cudaEventCreate(&p.b);
    cudaEventRecord(p.a, c->stream);
    v.push_back(p);
}
void prof_end(b200adsb_ctx *c, std::vector<EventPair> &v)
{
    if (!c->profile)
        return;
    cudaEventRecord(v.back().b, c->stream);
}
void prof_collect(b200adsb_ctx *c)   // after a stream sync
{
    for (auto &p : c->scan_events) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess)
            c->timing.scan_ms += ms;
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    for (auto &p : c->other_events) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess)
            c->timing.resolve_ms += ms;
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    c->scan_events.clear();
    c->other_events.clear();
}

int read_counters(b200adsb_ctx *c)
{
    CK(c, cudaMemcpyAsync(c->h_counters, c->d_counters, C_WORDS * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    prof_collect(c);
    return B200ADSB_OK;
}

// launch the scan kernel over buffers [b0, b0+nb) of the pending batch
int launch_scan(b200adsb_ctx *c, uint32_t b0, uint32_t nb)
{
    const Pending &q = c->cur;
    Scan7Smem L7(q.T);
    if (c->lut_T != q.T || c->lut_WP != L7.WP) {
        // field r of try-phase 4+tt of a candidate whose A = j+19 has A % 12 == ra starts at
        // 1/5-sample position 5*(A+e5)+z: word offset of its plane/residue stream in
        // plane[phi][rho][WP], and whether the stream index q = A/12 advances by one
        for (int i = 0; i < kLutWords; i++) {
            const int ra = i / 25, tt = (i - 25 * ra) / 5, r = i - 25 * ra - 5 * tt;
            const int e5 = (tt >= 1) ? 1 : 0, phi0 = (tt >= 1) ? tt - 1 : 4;
            const int z = phi0 + 12 * r, zd = z / 5, phi = z - 5 * zd;
            const int rr = ra + e5 + zd, wrap = rr >= 12 ? 1 : 0;
            c->h_lut[i] = (uint32_t)((phi * 12 + rr - 12 * wrap) * L7.WP) | ((uint32_t)wrap << 16);
        }
        CK(c, cudaStreamSynchronize(c->stream));   // a previous launch may still read the table
        CK(c, cudaMemcpyAsync(c->d_lut, c->h_lut, sizeof c->h_lut, cudaMemcpyHostToDevice, c->stream));
        c->lut_T = q.T;
        c->lut_WP = L7.WP;
    }
    Scan7Params P7{};
    ScanParams &p = P7.s;
    p.in = q.in;               // the kernel indexes buffers and tiles of the whole batch (carry mode reaches
    p.lengths = q.lengths;     // back into buffer b0-1 of the batch, not into the previous batch)
    p.b0 = b0;
    p.n_buffers = nb;
    p.spb = q.spb;
    p.stride = q.stride;
    p.T = q.T;
    p.tiles_per_buffer = q.tpb;
    p.vec_ok = (!q.from_mag && ((uintptr_t)q.in % 16 == 0) && (q.stride % 4 == 0)) ? 1 : 0;
    p.rec = c->d_rec;
    p.pool_cap = (uint32_t)std::min<size_t>(c->pool_cap, 0xffffffffu);
    p.tile_dir = c->d_tile_dir;
    p.counters = c->d_counters;
    p.ev_keys = c->d_ev_keys;
    p.ev_ord = c->d_ev_ord;
    p.ev_used = c->d_ev_used;
    p.ev_mask = kEvSlots - 1;
    p.ord_first = q.ord_first;
    p.ord_stride = q.ord_stride;
    p.crc_lanes = c->d_crc_lanes;
    p.lut = c->d_lut;
    p.carry = (c->carry && !q.from_mag && !q.msgs) ? 1 : 0;
    p.tails = c->d_tails;
    P7.off_planes = (uint32_t)L7.off_planes;
    P7.off_surv = (uint32_t)L7.off_surv;
    P7.off_masks = (uint32_t)L7.off_masks;
    P7.off_list = (uint32_t)L7.off_list;
    P7.off_cand = (uint32_t)L7.off_cand;
    P7.NG = L7.NG;
    P7.WP = L7.WP;
    P7.nw = L7.nw;
    P7.list_cap = L7.list_cap;
    P7.Wrow = (L7.NG + 1) / 2;
    P7.inv_Wrow = (65536 + P7.Wrow - 1) / P7.Wrow;
    const uint32_t grid = nb * (uint32_t)q.tpb;
    if (grid == 0)
        return B200ADSB_OK;
    prof_begin(c, c->scan_events);
    // the compile-time forms: default tile (shared memory plan as immediates), and on top of it a batch of
    // whole standard buffers (length, tiles per buffer, alignment and carry tests folded)
    const bool std_batch = !q.from_mag && q.T == kDefaultTile && !q.lengths && q.spb == (uint32_t)kMaxSamples &&
                           p.vec_ok && !p.carry;
    if (q.from_mag)
        scan7_kernel<true, 0, false><<<grid, k7Threads, L7.bytes, c->stream>>>(P7);
#ifndef B200_SCAN7_SPECIALISE
#define B200_SCAN7_SPECIALISE 2     // A/B builds: 0 = generic kernel only, 1 = + fixed tile, 2 = + standard batch
#endif
    else if (B200_SCAN7_SPECIALISE >= 2 && std_batch)
        scan7_kernel<false, kDefaultTile, true><<<grid, k7Threads, L7.bytes, c->stream>>>(P7);
    else if (B200_SCAN7_SPECIALISE >= 1 && q.T == kDefaultTile)
        scan7_kernel<false, kDefaultTile, false><<<grid, k7Threads, L7.bytes, c->stream>>>(P7);
    else
        scan7_kernel<false, 0, false><<<grid, k7Threads, L7.bytes, c->stream>>>(P7);
    prof_end(c, c->scan_events);
    CK(c, cudaGetLastError());
    c->timing.scan_launches++;
    c->timing.samples += (uint64_t)nb * q.spb;
    return B200ADSB_OK;
}

// per-batch counters to zero before stage 1.  A synchronous call also acknowledges a failed
// enqueue-only batch (it is the documented way to redo one): the sticky flag is cleared.
int reset_scan_counters(b200adsb_ctx *c, bool enqueue_only = false)
{
    if (!enqueue_only)
        CK(c, cudaMemsetAsync(c->d_counters + C_STICKY, 0, 4, c->stream));
    if (c->counters_clean) {   // the previous (enqueue-only) batch cleared them on the device
        c->counters_clean = false;
        return B200ADSB_OK;
    }
    // C_POOL, C_FLAGS cleared; C_CAND cleared; filter-full flag is per batch too
    CK(c, cudaMemsetAsync(c->d_counters + C_POOL, 0, 8, c->stream));
    CK(c, cudaMemsetAsync(c->d_counters + C_CAND, 0, 4, c->stream));
    return B200ADSB_OK;
}

int clear_events(b200adsb_ctx *c)
{
    // (C_FLAGS still holds the overflow that brought us here: nothing is admitted)
    const CommitArgs ca{c->d_ev_keys, c->d_ev_ord, c->d_ev_used, c->d_new_keys, c->d_counters, c->d_members,
                        nullptr, 0u, 0};
    events_commit_kernel<<<1, 1024, 0, c->stream>>>(ca);
    CK(c, cudaGetLastError());
    c->timing.other_launches++;
    return B200ADSB_OK;
}

// set up a pending batch and run stage 1 (with pool growth); on return the scan is complete
int scan_begin(b200adsb_ctx *c, const void *d_in, bool from_mag, size_t n_buffers, size_t spb,
               size_t stride, const uint32_t *d_lengths, unsigned long long ord_first,
               unsigned long long ord_stride)
{
    if (c->cur.active)
        return B200ADSB_ERR_STATE;
    if (spb > (size_t)kMaxSamples || n_buffers > 0xffffffu)
        return B200ADSB_ERR_BAD_ARG;
    Pending &q = c->cur;
    q = Pending();
    q.in = d_in;
    q.from_mag = from_mag;
    q.lengths = d_lengths;
    q.n_buffers = (uint32_t)n_buffers;
    q.spb = (uint32_t)spb;
    q.stride = stride;
    q.T = pick_tile(c, n_buffers, spb);
    q.tpb = spb ? (int)((spb + q.T - 1) / q.T) : 0;
    q.n_tiles = (uint32_t)(n_buffers * (size_t)q.tpb);
    q.ord_first = ord_first;
    q.ord_stride = ord_stride;
    int rc = ensure_tiles(c, std::max<size_t>(q.n_tiles, 1));
    if (rc) return rc;
    const size_t positions = n_buffers * spb;
    rc = ensure_pool(c, std::max<size_t>(positions >> c->pool_shift, 4096));
    if (rc) return rc;
    q.active = true;
    return B200ADSB_OK;
}

// runs/re-runs stage 1 over the whole pending batch until the pool suffices
int scan_run_all(b200adsb_ctx *c)
{
    Pending &q = c->cur;
    for (int attempt = 0; attempt < 8; attempt++) {
        int rc = reset_scan_counters(c);
        if (rc) return rc;
        rc = launch_scan(c, 0, q.n_buffers);
        if (rc) return rc;
        rc = read_counters(c);
        if (rc) return rc;
        const uint32_t flags = c->h_counters[C_FLAGS];
        if (flags & F_EV_OVF) {
            clear_events(c);
            q.active = false;
            return B200ADSB_ERR_EVENTS;
        }
        if (!(flags & F_POOL_OVF))
            return B200ADSB_OK;
        // candidate pool too small: recycle the partial events, grow, redo
        rc = clear_events(c);
        if (rc) return rc;
        const size_t need = (size_t)c->h_counters[C_POOL];
        rc = ensure_pool(c, need + need / 8 + 1024);
        if (rc) return rc;
    }
    q.active = false;
    return B200ADSB_ERR_NOMEM;
}

// every exit of a call that owns a pending batch ends it (a failed call must not leave the
// context refusing all later calls with ERR_STATE)
struct PendingGuard {
    b200adsb_ctx *c;
    ~PendingGuard() { c->cur.active = false; }
};

// stage 2: finalise events, resolve, ordered emit, commit
int resolve_run_impl(b200adsb_ctx *c, b200adsb_frame *d_out, size_t cap, size_t *n_out,
                     uint32_t *d_per_buffer_counts, uint32_t *d_async_result);
int resolve_run(b200adsb_ctx *c, b200adsb_frame *d_out, size_t cap, size_t *n_out,
                uint32_t *d_per_buffer_counts, uint32_t *d_async_result = nullptr)
{
    const int rc = resolve_run_impl(c, d_out, cap, n_out, d_per_buffer_counts, d_async_result);
    if (rc != kRedo)      // kRedo: the caller redoes the scan on the same pending batch
        c->cur.active = false;
    return rc;
}
int resolve_run_impl(b200adsb_ctx *c, b200adsb_frame *d_out, size_t cap, size_t *n_out,
                     uint32_t *d_per_buffer_counts, uint32_t *d_async_result)
{
    Pending &q = c->cur;
    if (!q.active)
        return B200ADSB_ERR_STATE;
    prof_begin(c, c->other_events);
    const uint32_t n_ctas = (q.n_tiles + 31) / 32;
    FinalizeArgs fa{c->d_ev_keys, c->d_ev_ord, c->d_ev_used, c->d_ev_tmp, c->d_new_keys, c->d_counters,
                    c->d_members, kEvSlots - 1, c->d_bloom};
    ResolveParams rp{};
    rp.rec = c->d_rec;
    rp.tile_dir = c->d_tile_dir;
    rp.n_tiles = q.n_tiles;
    rp.tiles_per_buffer = std::max(q.tpb, 1);
    rp.emit_info = c->d_emit_info;
    rp.tile_emit = c->d_tile_emit;
    rp.cta_sum = c->d_cta_sum;
    rp.bloom = c->d_bloom;
    rp.rec_score = q.msgs ? c->d_rec_score : nullptr;   // only the message-level API reads scores
    rp.members = c->d_members;
    rp.ev_keys = c->d_ev_keys;
    rp.ev_ord = c->d_ev_ord;
    rp.ev_mask = kEvSlots - 1;
    rp.ord_first = q.ord_first;
    rp.ord_stride = q.ord_stride;
    EmitParams ep{};
    ep.in = q.in;
    ep.lengths = q.lengths;
    ep.spb = q.spb;
    ep.stride = q.stride;
    ep.rec = c->d_rec;
    ep.tile_dir = c->d_tile_dir;
    ep.emit_info = c->d_emit_info;
    ep.carry = (c->carry && !q.from_mag && !q.msgs) ? 1 : 0;
    ep.tails = c->d_tails;
    ep.counters = c->d_counters;
    ep.tile_cnt = c->d_tile_emit;
    ep.cta_excl = c->d_cta_sum;
    ep.n_tiles = q.n_tiles;
    ep.tiles_per_buffer = std::max(q.tpb, 1);
    ep.out = d_out;
    ep.cap = (uint32_t)std::min<size_t>(cap, 0xffffffffu);
    ep.msgs = q.msgs;
    const bool save_tail = c->carry && !q.from_mag && !q.msgs && q.n_buffers > 0;
    const bool small = n_ctas <= 16 && !d_per_buffer_counts && !d_async_result;
    if (small) {
        // one launch for the whole second stage (the tail is saved first: commit ends the kernel)
        if (save_tail) {
            save_tail_kernel<<<1, kTailWords, 0, c->stream>>>(reinterpret_cast<const uint32_t *>(q.in), q.stride,
                                                              q.lengths, q.spb, q.n_buffers, c->d_tails, c->d_counters);
            CK(c, cudaGetLastError());
        }
        if (q.from_mag)
            resolve_small_kernel<true><<<1, kResolveThreads, 0, c->stream>>>(fa, rp, ep, n_ctas, save_tail ? 1 : 0);
        else
            resolve_small_kernel<false><<<1, kResolveThreads, 0, c->stream>>>(fa, rp, ep, n_ctas, save_tail ? 1 : 0);
        CK(c, cudaGetLastError());
        c->timing.other_launches += 1 + (save_tail ? 1 : 0);
    } else {
        events_finalize_kernel<<<1, 1024, 0, c->stream>>>(fa);
        CK(c, cudaGetLastError());
        if (q.n_tiles) {
            // the last block also scans the per-block sums (32 tiles each); total -> counters[C_FRAMES]
            resolve_kernel<<<n_ctas, kResolveThreads, 0, c->stream>>>(rp, c->d_counters);
            CK(c, cudaGetLastError());
        } else {
            tile_scan_kernel<<<1, 1024, 0, c->stream>>>(c->d_cta_sum, n_ctas, c->d_counters);
            CK(c, cudaGetLastError());
        }
        if (d_per_buffer_counts && q.n_buffers) {
            buffer_counts_kernel<<<(q.n_buffers + 255) / 256, 256, 0, c->stream>>>(
                c->d_tile_emit, q.n_buffers, q.tpb, d_per_buffer_counts);
            CK(c, cudaGetLastError());
            c->timing.other_launches++;
        }
        const CommitArgs ca{c->d_ev_keys, c->d_ev_ord, c->d_ev_used, c->d_new_keys, c->d_counters, c->d_members,
                            d_async_result, ep.cap, save_tail ? 1 : 0};
        const bool fused_commit = q.n_tiles && !save_tail;   // (the tail is saved before the commit)
        if (q.n_tiles) {
            // one warp per frame; the frame count is only known on the device, so a fixed grid strides;
            // one extra block runs the commit step
            const uint32_t eg = (uint32_t)std::min<size_t>(148 * 8, std::max<size_t>(1, (cap + kEmitWarps - 1) / kEmitWarps));
            CommitArgs none{};
            if (q.from_mag)
                emit_frames_kernel<true><<<eg + 1, 32 * kEmitWarps, 0, c->stream>>>(ep, c->d_counters, n_ctas, fused_commit ? ca : none);
            else
                emit_frames_kernel<false><<<eg + 1, 32 * kEmitWarps, 0, c->stream>>>(ep, c->d_counters, n_ctas, fused_commit ? ca : none);
            CK(c, cudaGetLastError());
        }
        if (save_tail) {
            save_tail_kernel<<<1, kTailWords, 0, c->stream>>>(reinterpret_cast<const uint32_t *>(q.in), q.stride,
                                                              q.lengths, q.spb, q.n_buffers, c->d_tails, c->d_counters);
            CK(c, cudaGetLastError());
        }
        if (!fused_commit) {
            events_commit_kernel<<<1, 1024, 0, c->stream>>>(ca);
            CK(c, cudaGetLastError());
            c->timing.other_launches++;
        }
        c->timing.other_launches += 3;
    }
    if (d_async_result) {
        // enqueue-only form: the batch outcome stays on the device (written by the commit kernel, which
        // also cleared the per-batch counters), nothing is read back here
        c->counters_clean = !small;
        prof_end(c, c->other_events);
        q.active = false;
        return B200ADSB_OK;
    }
    prof_end(c, c->other_events);
    int rc = read_counters(c);
    if (rc) { q.active = false; return rc; }
    if (c->h_counters[C_FLAGS] & kBadMask)
        return kRedo;   // the batch failed (here or on another rank): nothing was committed, the caller redoes it
    q.active = false;
    c->timing.candidates += c->h_counters[C_CAND];
    const size_t n = c->h_counters[C_FRAMES];
    if (n_out)
        *n_out = n;
    return n > cap ? B200ADSB_ERR_CAPACITY : B200ADSB_OK;
}

// stage 1 without a host round trip + stage 2; if the scan turns out to have overflowed the
// candidate pool (rare: the pool is sized at ~3x the typical survivor rate and grows), redo it
// with the checked path.  `scanned` = stage 1 has already been launched by the caller.
int scan_resolve(b200adsb_ctx *c, bool scanned, b200adsb_frame *d_out, size_t cap, size_t *n_out,
                 uint32_t *d_per_buffer_counts)
{
    PendingGuard guard{c};
    int rc;
    if (!scanned) {
        rc = reset_scan_counters(c);
        if (rc) return rc;
        rc = launch_scan(c, 0, c->cur.n_buffers);
        if (rc) { c->cur.active = false; return rc; }
    }
    rc = resolve_run(c, d_out, cap, n_out, d_per_buffer_counts);
    if (rc != kRedo)
        return rc;
    if (c->h_counters[C_FLAGS] & F_EV_OVF) {
        c->cur.active = false;
        return B200ADSB_ERR_EVENTS;
    }
    const size_t need = (size_t)c->h_counters[C_POOL];
    rc = ensure_pool(c, need + need / 8 + 1024);
    if (rc) { c->cur.active = false; return rc; }
    rc = scan_run_all(c);
    if (rc) { c->cur.active = false; return rc; }
    rc = resolve_run(c, d_out, cap, n_out, d_per_buffer_counts);
    return rc == kRedo ? B200ADSB_ERR_NOMEM : rc;
}

int ensure_stage(b200adsb_ctx *c, size_t bytes)
{
    return grow(c, (unsigned char **)&c->d_stage, &c->stage_bytes, bytes, 1);
}

}  // namespace

// =================================================================== C ABI
extern "C" {

int b200adsb_version(void) { return 100; }

const char *b200adsb_strerror(int s)
{
    switch (s) {
    case B200ADSB_OK: return "ok";
    case B200ADSB_ERR_BAD_ARG: return "bad argument";
    case B200ADSB_ERR_CAPACITY: return "output capacity exceeded";
    case B200ADSB_ERR_CUDA: return "CUDA error";
    case B200ADSB_ERR_NOMEM: return "out of device memory";
    case B200ADSB_ERR_STATE: return "call out of order";
    case B200ADSB_ERR_EVENTS: return "too many distinct new ICAO addresses in one batch";
    default: return "unknown status";
    }
}

const char *b200adsb_last_error(const b200adsb_ctx *ctx) { return ctx ? ctx->err : "null context"; }

int b200adsb_ctx_create(b200adsb_ctx **out, int device, void *stream)
{
    if (!out)
        return B200ADSB_ERR_BAD_ARG;
    *out = nullptr;
    b200adsb_ctx *c = new (std::nothrow) b200adsb_ctx();
    if (!c)
        return B200ADSB_ERR_NOMEM;
    c->device = device;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        delete c;
        return B200ADSB_ERR_CUDA;   // no CPU fallback: without a GPU there is no context
    }
    auto fail = [&](int rc) {
        b200adsb_ctx_destroy(c);
        return rc;
    };
#define CKC(call)                          \
    do {                                   \
        if ((call) != cudaSuccess)         \
            return fail(B200ADSB_ERR_CUDA); \
    } while (0)
    CKC(cudaSetDevice(device));
    {   // the scan kernel's launch attributes, once: room for the largest tile, all of L1 as shared memory
        const int max_smem = (int)Scan7Smem(kMaxTile).bytes;
        auto prep = [&](const void *fn) {
            return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem) == cudaSuccess &&
                   cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, 100) == cudaSuccess;
        };
        if (!prep((const void *)scan7_kernel<true, 0, false>) || !prep((const void *)scan7_kernel<false, 0, false>) ||
            !prep((const void *)scan7_kernel<false, kDefaultTile, false>) ||
            !prep((const void *)scan7_kernel<false, kDefaultTile, true>))
            return fail(B200ADSB_ERR_CUDA);
    }
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        CKC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    CKC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CKC(cudaMalloc((void **)&c->d_counters, C_WORDS * 4));
    CKC(cudaMemset(c->d_counters, 0, C_WORDS * 4));
    CKC(cudaMallocHost((void **)&c->h_counters, C_WORDS * 4));
    CKC(cudaMalloc((void **)&c->d_members, kMemberSlots * 4));
    CKC(cudaMemset(c->d_members, 0, kMemberSlots * 4));
    CKC(cudaMalloc((void **)&c->d_ev_keys, kEvSlots * 4));
    CKC(cudaMemset(c->d_ev_keys, 0, kEvSlots * 4));
    CKC(cudaMalloc((void **)&c->d_ev_ord, kEvSlots * 8));
    CKC(cudaMemset(c->d_ev_ord, 0xff, kEvSlots * 8));
    CKC(cudaMalloc((void **)&c->d_ev_used, kEvSlots * 4));
    CKC(cudaMalloc((void **)&c->d_ev_tmp, kEvSlots * 4));
    CKC(cudaMalloc((void **)&c->d_new_keys, (size_t)B200ADSB_ICAO_FILTER_SIZE * 4));
    CKC(cudaMalloc((void **)&c->d_crc256, 256 * 4));
    CKC(cudaMalloc((void **)&c->d_crc_lanes, kLaneTabs * 32 * 4));
    CKC(cudaMalloc((void **)&c->d_lut, kLutWords * 4));
    CKC(cudaMalloc((void **)&c->d_bloom, kBloomWords * 4));
    CKC(cudaMalloc((void **)&c->d_tails, 2 * kTailWords * 4));
    CKC(cudaMemset(c->d_tails, 0, 2 * kTailWords * 4));
    CKC(cudaMalloc((void **)&c->d_scalar, 64));
    {
        uint32_t t[kTabWords], t256[256], tl[kLaneTabs * 32];
        build_crc_tabs(t, t256, tl);
        CKC(cudaMemcpy(c->d_crc_lanes, tl, sizeof tl, cudaMemcpyHostToDevice));
        CKC(cudaMemcpy(c->d_crc256, t256, sizeof t256, cudaMemcpyHostToDevice));
    }
#undef CKC
    *out = c;
    return B200ADSB_OK;
}

void b200adsb_ctx_destroy(b200adsb_ctx *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    if (c->stream)
        cudaStreamSynchronize(c->stream);
    prof_collect(c);
    for (auto e : c->chunk_events)
        cudaEventDestroy(e);
    cudaFree(c->d_counters);
    if (c->h_counters) cudaFreeHost(c->h_counters);
    cudaFree(c->d_members);
    cudaFree(c->d_ev_keys);
    cudaFree(c->d_ev_ord);
    cudaFree(c->d_ev_used);
    cudaFree(c->d_ev_tmp);
    cudaFree(c->d_new_keys);
    cudaFree(c->d_crc256);
    cudaFree(c->d_crc_lanes);
    cudaFree(c->d_lut);
    cudaFree(c->d_scalar);
    cudaFree(c->d_rec);
    cudaFree(c->d_emit_info);
    cudaFree(c->d_rec_score);
    cudaFree(c->d_tile_dir);
    cudaFree(c->d_tile_emit);
    cudaFree(c->d_cta_sum);
    cudaFree(c->d_bloom);
    cudaFree(c->d_tails);
    for (auto &sl : c->slots) {
        cudaFree(sl.d_iq);
        cudaFree(sl.d_lengths);
        cudaFree(sl.d_frames);
        cudaFree(sl.d_result);
        if (sl.done) cudaEventDestroy(sl.done);
    }
    cudaFree(c->d_stage);
    cudaFree(c->d_stage8);
    cudaFree(c->d_frames);
    cudaFree(c->d_counts);
    cudaFree(c->d_lengths);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int b200adsb_ctx_set_option(b200adsb_ctx *c, int option, int64_t value)
{
    if (!c)
        return B200ADSB_ERR_BAD_ARG;
    switch (option) {
    case B200ADSB_OPT_TILE:
        if (value != 0 && (value < 8 || value > kMaxTile || value % 8))
            return B200ADSB_ERR_BAD_ARG;
        c->tile_opt = (int)value;
        return B200ADSB_OK;
    case B200ADSB_OPT_POOL_SHIFT:
        if (value < 0 || value > 16)
            return B200ADSB_ERR_BAD_ARG;
        c->pool_shift = (int)value;
        return B200ADSB_OK;
    case B200ADSB_OPT_PROFILE:
        c->profile = value ? 1 : 0;
        return B200ADSB_OK;
    case B200ADSB_OPT_CARRY:
        c->carry = value ? 1 : 0;   // (re)starting continuity: nothing precedes the next buffer
        if (cudaSetDevice(c->device) != cudaSuccess ||
            cudaMemsetAsync(c->d_tails, 0, 2 * kTailWords * 4, c->stream) != cudaSuccess)
            return B200ADSB_ERR_CUDA;
        return B200ADSB_OK;
    case B200ADSB_OPT_H2D_CHUNK:
        if (value < 1)
            return B200ADSB_ERR_BAD_ARG;
        c->h2d_chunk = (int)value;
        return B200ADSB_OK;
    default:
        return B200ADSB_ERR_BAD_ARG;
    }
}

int b200adsb_ctx_sync(b200adsb_ctx *c)
{
    if (!c)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    CK(c, cudaStreamSynchronize(c->stream));
    prof_collect(c);
    return B200ADSB_OK;
}

int b200adsb_timing_get(b200adsb_ctx *c, b200adsb_timing *out, int reset)
{
    if (!c || !out)
        return B200ADSB_ERR_BAD_ARG;
    *out = c->timing;
    if (reset)
        c->timing = b200adsb_timing{};
    return B200ADSB_OK;
}

void *b200adsb_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess)
        return nullptr;
    return p;
}
void b200adsb_host_free(void *p)
{
    if (p)
        cudaFreeHost(p);
}

// ------------------------------------------------------------------ to_mag
int b200adsb_to_mag(b200adsb_ctx *c, const int16_t *iq, size_t n, uint16_t *data, size_t *length)
{
    if (!c || (!iq && n) || !data)
        return B200ADSB_ERR_BAD_ARG;
    if (n > (size_t)kMaxSamples)
        return B200ADSB_ERR_BAD_ARG;   // reference: index out of bounds panic, lib.rs:48
    int rc = bind(c);
    if (rc) return rc;
    rc = ensure_stage(c, (size_t)kMaxSamples * 4 + (size_t)kMagLen * 2 + 64);
    if (rc) return rc;
    uint32_t *d_iq = reinterpret_cast<uint32_t *>(c->d_stage);
    uint16_t *d_mag = reinterpret_cast<uint16_t *>(reinterpret_cast<unsigned char *>(c->d_stage) + (size_t)kMaxSamples * 4);
    if (n)
        CK(c, cudaMemcpyAsync(d_iq, iq, n * 4, cudaMemcpyHostToDevice, c->stream));
    to_mag_kernel<<<(kMagLen + 255) / 256, 256, 0, c->stream>>>(d_iq, (int)n, d_mag);
    CK(c, cudaGetLastError());
    c->timing.other_launches++;
    CK(c, cudaMemcpyAsync(data, d_mag, (size_t)kMagLen * 2, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    if (length)
        *length = n;
    return B200ADSB_OK;
}

// ------------------------------------------------------------------ device batch
int b200adsb_scan_batch_dev(b200adsb_ctx *c, const int16_t *d_iq, size_t n_buffers, size_t spb,
                            size_t stride, const uint32_t *d_lengths, uint64_t first_ordinal,
                            uint64_t ordinal_stride)
{
    if (!c || (!d_iq && n_buffers && spb))
        return B200ADSB_ERR_BAD_ARG;
    if (stride < spb && n_buffers > 1)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    rc = scan_begin(c, d_iq, false, n_buffers, spb, stride, d_lengths, first_ordinal,
                    ordinal_stride ? ordinal_stride : 1);
    if (rc) return rc;
    rc = scan_run_all(c);
    if (rc)
        c->cur.active = false;
    return rc;
}

// enqueue-only forms of the sharded pair (see b200adsb_demod_iq_batch_dev_async)
int b200adsb_scan_batch_dev_async(b200adsb_ctx *c, const int16_t *d_iq, size_t n_buffers, size_t spb,
                                  size_t stride, const uint32_t *d_lengths, uint64_t first_ordinal,
                                  uint64_t ordinal_stride)
{
    if (!c || (!d_iq && n_buffers && spb) || (stride < spb && n_buffers > 1))
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    rc = scan_begin(c, d_iq, false, n_buffers, spb, stride, d_lengths, first_ordinal,
                    ordinal_stride ? ordinal_stride : 1);
    if (rc) return rc;
    rc = reset_scan_counters(c, true);
    if (!rc)
        rc = launch_scan(c, 0, c->cur.n_buffers);
    if (rc)
        c->cur.active = false;
    return rc;
}

int b200adsb_resolve_batch_dev_async(b200adsb_ctx *c, b200adsb_frame *d_out, size_t cap, uint32_t *d_result)
{
    if (!c || (!d_out && cap) || !d_result)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    return resolve_run(c, d_out, cap, nullptr, nullptr, d_result);
}

int b200adsb_events_count(b200adsb_ctx *c, size_t *n)
{
    if (!c || !n)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    rc = read_counters(c);
    if (rc) return rc;
    *n = c->h_counters[C_EV_USED];
    return B200ADSB_OK;
}

int b200adsb_events_export_dev(b200adsb_ctx *c, uint64_t *d_pairs, size_t cap, size_t *n)
{
    if (!c || (!d_pairs && cap))
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    events_export_kernel<<<32, 256, 0, c->stream>>>(c->d_ev_keys, c->d_ev_ord, c->d_ev_used, c->d_counters,
                                                    (unsigned long long *)d_pairs,
                                                    (uint32_t)std::min<size_t>(cap, 0xffffffffu));
    CK(c, cudaGetLastError());
    c->timing.other_launches++;
    rc = read_counters(c);
    if (rc) return rc;
    const size_t have = c->h_counters[C_EV_USED];
    if (n)
        *n = have;
    return have > cap ? B200ADSB_ERR_CAPACITY : B200ADSB_OK;
}

int b200adsb_events_import_dev(b200adsb_ctx *c, const uint64_t *d_pairs, size_t n)
{
    if (!c || (!d_pairs && n))
        return B200ADSB_ERR_BAD_ARG;
    if (!n)
        return B200ADSB_OK;
    int rc = bind(c);
    if (rc) return rc;
    events_import_kernel<<<32, 256, 0, c->stream>>>((const unsigned long long *)d_pairs, (uint32_t)n,
                                                    c->d_ev_keys, c->d_ev_ord, c->d_ev_used,
                                                    kEvSlots - 1, c->d_counters);
    CK(c, cudaGetLastError());
    c->timing.other_launches++;
    return B200ADSB_OK;
}

int b200adsb_events_pack_dev(b200adsb_ctx *c, uint64_t *d_rows, size_t rows_cap)
{
    if (!c || !d_rows || rows_cap < 2)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    events_pack_kernel<<<16, 256, 0, c->stream>>>(c->d_ev_keys, c->d_ev_ord, c->d_ev_used, c->d_counters,
                                                  (unsigned long long *)d_rows,
                                                  (uint32_t)std::min<size_t>(rows_cap, 0xffffffffu));
    CK(c, cudaGetLastError());
    c->timing.other_launches++;
    return B200ADSB_OK;
}

int b200adsb_events_import_packed_dev(b200adsb_ctx *c, const uint64_t *d_gathered, size_t n_ranks,
                                      size_t rows_per_rank, size_t skip_rank)
{
    if (!c || !d_gathered || rows_per_rank < 2)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    events_import_packed_kernel<<<16, 256, 0, c->stream>>>(
        (const unsigned long long *)d_gathered, (uint32_t)n_ranks, (uint32_t)rows_per_rank, (uint32_t)skip_rank,
        c->d_ev_keys, c->d_ev_ord, c->d_ev_used, kEvSlots - 1, c->d_counters);
    CK(c, cudaGetLastError());
    c->timing.other_launches++;
    return B200ADSB_OK;
}

size_t b200adsb_events_symm_words(size_t n_ranks, size_t rows_per_rank)
{
    return 2 * n_ranks + 2 * n_ranks * rows_per_rank * 2;
}

int b200adsb_events_push_symm_dev(b200adsb_ctx *c, uint64_t *const *d_peer_bufs, size_t rank, size_t n_ranks,
                                  size_t rows_per_rank, uint64_t epoch, uint32_t force_flags)
{
    if (!c || !d_peer_bufs || n_ranks == 0 || n_ranks > 256 || rank >= n_ranks || rows_per_rank < 2 || epoch == 0)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    events_push_symm_kernel<<<16, 256, 0, c->stream>>>(c->d_ev_keys, c->d_ev_ord, c->d_ev_used, c->d_counters,
                                                        (unsigned long long *const *)d_peer_bufs, (uint32_t)rank,
                                                        (uint32_t)n_ranks, (uint32_t)rows_per_rank, epoch,
                                                        force_flags & kBadMask);
    CK(c, cudaGetLastError());
    c->timing.other_launches++;
    return B200ADSB_OK;
}

int b200adsb_events_import_symm_dev(b200adsb_ctx *c, uint64_t *d_local_buf, size_t rank, size_t n_ranks,
                                    size_t rows_per_rank, uint64_t epoch)
{
    if (!c || !d_local_buf || n_ranks == 0 || n_ranks > 256 || rank >= n_ranks || rows_per_rank < 2 || epoch == 0)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    events_import_symm_kernel<<<16, 256, 0, c->stream>>>((unsigned long long *)d_local_buf, (uint32_t)rank,
                                                          (uint32_t)n_ranks, (uint32_t)rows_per_rank, epoch,
                                                          c->d_ev_keys, c->d_ev_ord, c->d_ev_used, kEvSlots - 1,
                                                          c->d_counters);
    CK(c, cudaGetLastError());
    c->timing.other_launches++;
    return B200ADSB_OK;
}

int b200adsb_frames_pack_dev(b200adsb_ctx *c, void *stream, const b200adsb_frame *d_frames, const uint32_t *d_count,
                             size_t count, b200adsb_frame *d_block, size_t rows_cap)
{
    cudaStream_t st = stream ? (cudaStream_t)stream : c ? c->stream : nullptr;
    if (!c || !d_block || rows_cap == 0 || rows_cap > 0xfffffffeu || (!d_frames && (d_count || count)))
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    frames_pack_kernel<<<8, 256, 0, st>>>(d_frames, d_count, (uint32_t)std::min<size_t>(count, 0xffffffffu),
                                                  d_block, (uint32_t)rows_cap);
    CK(c, cudaGetLastError());
    c->timing.other_launches++;
    return B200ADSB_OK;
}

int b200adsb_frames_merge_dev(b200adsb_ctx *c, void *stream, const b200adsb_frame *d_gathered, size_t n_ranks,
                              size_t rows_cap, b200adsb_frame *d_out, size_t cap, uint32_t *d_n_out)
{
    cudaStream_t st = stream ? (cudaStream_t)stream : c ? c->stream : nullptr;
    if (!c || !d_gathered || !d_n_out || n_ranks == 0 || rows_cap == 0 || rows_cap > 0xfffffffeu || (!d_out && cap))
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    frames_merge_kernel<<<16, 256, 0, st>>>(d_gathered, (uint32_t)n_ranks, (uint32_t)rows_cap, d_out,
                                                   (uint32_t)std::min<size_t>(cap, 0xffffffffu), d_n_out, nullptr, 0ull);
    CK(c, cudaGetLastError());
    c->timing.other_launches++;
    return B200ADSB_OK;
}

size_t b200adsb_frames_symm_bytes(size_t n_ranks, size_t rows_cap)
{
    return 16 * n_ranks + 2 * n_ranks * (rows_cap + 1) * sizeof(b200adsb_frame);
}

int b200adsb_frames_push_symm_dev(b200adsb_ctx *c, void *stream, const b200adsb_frame *d_frames,
                                  const uint32_t *d_count, size_t count, void *const *d_peer_bufs, size_t rank,
                                  size_t n_ranks, size_t rows_cap, uint64_t epoch, uint32_t *d_ticket)
{
    if (!c || !d_peer_bufs || !d_ticket || n_ranks == 0 || n_ranks > 256 || rank >= n_ranks || rows_cap == 0 ||
        rows_cap > 0xfffffffeu || epoch == 0 || (!d_frames && (d_count || count)))
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    frames_push_symm_kernel<<<8, 256, 0, st>>>(d_frames, d_count, (uint32_t)std::min<size_t>(count, 0xffffffffu),
                                               (unsigned char *const *)d_peer_bufs, (uint32_t)rank, (uint32_t)n_ranks,
                                               (uint32_t)rows_cap, epoch, d_ticket);
    CK(c, cudaGetLastError());
    c->timing.other_launches++;
    return B200ADSB_OK;
}

int b200adsb_frames_merge_symm_dev(b200adsb_ctx *c, void *stream, void *d_local_buf, size_t n_ranks,
                                   size_t rows_cap, uint64_t epoch, b200adsb_frame *d_out, size_t cap,
                                   uint32_t *d_n_out)
{
    if (!c || !d_local_buf || !d_n_out || n_ranks == 0 || n_ranks > 256 || rows_cap == 0 || rows_cap > 0xfffffffeu ||
        epoch == 0 || (!d_out && cap))
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const uint32_t parity = (uint32_t)(epoch & 1);
    unsigned char *base = reinterpret_cast<unsigned char *>(d_local_buf);
    const b200adsb_frame *blocks = reinterpret_cast<const b200adsb_frame *>(base + 16 * n_ranks) +
                                   (size_t)parity * n_ranks * (rows_cap + 1);
    const unsigned long long *flags = reinterpret_cast<const unsigned long long *>(base) + (size_t)parity * n_ranks;
    frames_merge_kernel<<<16, 256, 0, st>>>(blocks, (uint32_t)n_ranks, (uint32_t)rows_cap, d_out,
                                            (uint32_t)std::min<size_t>(cap, 0xffffffffu), d_n_out, flags, epoch);
    CK(c, cudaGetLastError());
    c->timing.other_launches++;
    return B200ADSB_OK;
}

int b200adsb_resolve_batch_dev(b200adsb_ctx *c, b200adsb_frame *d_out, size_t cap, size_t *n_out,
                               uint32_t *d_per_buffer_counts)
{
    if (!c || (!d_out && cap))
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    rc = resolve_run(c, d_out, cap, n_out, d_per_buffer_counts);
    if (rc == kRedo) {   // only an exchange-buffer or event-table overflow can surface here
        c->cur.active = false;
        return B200ADSB_ERR_EVENTS;
    }
    return rc;
}

int b200adsb_demod_iq_batch_dev_async(b200adsb_ctx *c, const int16_t *d_iq, size_t n_buffers, size_t spb,
                                      size_t stride, const uint32_t *d_lengths, b200adsb_frame *d_out,
                                      size_t cap, uint32_t *d_result)
{
    if (!c || !d_result)
        return B200ADSB_ERR_BAD_ARG;
    if ((!d_iq && n_buffers && spb) || (stride < spb && n_buffers > 1))
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    rc = scan_begin(c, d_iq, false, n_buffers, spb, stride, d_lengths, c->next_ordinal, 1);
    if (rc) return rc;
    c->next_ordinal += n_buffers;
    rc = reset_scan_counters(c, true);
    if (rc) { c->cur.active = false; return rc; }
    rc = launch_scan(c, 0, c->cur.n_buffers);
    if (rc) { c->cur.active = false; return rc; }
    rc = resolve_run(c, d_out, cap, nullptr, nullptr, d_result);
    if (rc)
        c->cur.active = false;
    return rc;
}

int b200adsb_demod_iq_batch_dev(b200adsb_ctx *c, const int16_t *d_iq, size_t n_buffers, size_t spb,
                                size_t stride, const uint32_t *d_lengths, b200adsb_frame *d_out,
                                size_t cap, size_t *n_out, uint32_t *d_per_buffer_counts)
{
    if (!c)
        return B200ADSB_ERR_BAD_ARG;
    if ((!d_iq && n_buffers && spb) || (stride < spb && n_buffers > 1))
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    rc = scan_begin(c, d_iq, false, n_buffers, spb, stride, d_lengths, c->next_ordinal, 1);
    if (rc) return rc;
    c->next_ordinal += n_buffers;
    return scan_resolve(c, false, d_out, cap, n_out, d_per_buffer_counts);
}

// ------------------------------------------------------------------ host batch
}  // extern "C"
namespace {
// iq8 != nullptr: the source is 8-bit unsigned (I, Q) pairs (RTL-SDR native); they cross PCIe as they are
// (2 bytes per sample) and are expanded to CS16 on the device by cu8_expand_kernel
int host_batch(b200adsb_ctx *c, const int16_t *iq, const uint8_t *iq8, size_t n_buffers, size_t spb,
               size_t stride, const uint32_t *lengths, b200adsb_frame *out, size_t cap,
               size_t *n_out, uint32_t *per_buffer_counts)
{
    if (!c || (!iq && !iq8 && n_buffers && spb) || (!out && cap))
        return B200ADSB_ERR_BAD_ARG;
    if (spb > (size_t)kMaxSamples || (stride < spb && n_buffers > 1))
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    if (n_out)
        *n_out = 0;
    // device staging is dense (stride == spb rounded up to 4 samples for 16 B alignment)
    const size_t dstride = (spb + 3) & ~(size_t)3;
    const size_t per_pass = std::max<size_t>(1, dstride ? kMaxStageBytes / (dstride * 4 + 4) : n_buffers);
    size_t done = 0, total_frames = 0;
    int status = B200ADSB_OK;
    do {
        const size_t nb = std::min(per_pass, n_buffers - done);
        rc = ensure_stage(c, std::max<size_t>(nb * dstride * 4, 16));
        if (rc) return rc;
        int16_t *d_iq = reinterpret_cast<int16_t *>(c->d_stage);
        if (iq8) {
            rc = grow(c, &c->d_stage8, &c->stage8_bytes, std::max<size_t>(nb * dstride * 2, 16));
            if (rc) return rc;
        }
        const uint32_t *d_len = nullptr;
        if (lengths) {
            rc = grow(c, &c->d_lengths, &c->lengths_cap, nb);
            if (rc) return rc;
            CK(c, cudaMemcpyAsync(c->d_lengths, lengths + done, nb * 4, cudaMemcpyHostToDevice, c->stream));
            d_len = c->d_lengths;
        }
        const size_t room = cap > total_frames ? cap - total_frames : 0;
        rc = grow(c, &c->d_frames, &c->frames_cap, std::max<size_t>(room, 1));
        if (rc) return rc;
        if (per_buffer_counts) {
            rc = grow(c, &c->d_counts, &c->counts_cap, nb);
            if (rc) return rc;
        }
        rc = scan_begin(c, d_iq, false, nb, spb, dstride, d_len, c->next_ordinal, 1);
        if (rc) return rc;
        PendingGuard guard{c};
        // H2D in chunks on the copy stream, scan of chunk k overlapping the copy of chunk k+1
        const size_t chunk = (size_t)c->h2d_chunk;
        const size_t n_chunks = nb ? (nb + chunk - 1) / chunk : 0;
        while (c->chunk_events.size() < n_chunks) {
            cudaEvent_t e;
            CK(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            c->chunk_events.push_back(e);
        }
        rc = reset_scan_counters(c);
        if (rc) return rc;
        for (size_t k = 0; k < n_chunks; k++) {
            const size_t b0 = k * chunk, cb = std::min(chunk, nb - b0);
            const int16_t *src = iq + 2 * (done + b0) * stride;
            if (spb && iq8) {
                uint8_t *d8 = c->d_stage8 + 2 * b0 * dstride;
                const uint8_t *src8 = iq8 + 2 * (done + b0) * stride;
                if (stride == dstride)
                    CK(c, cudaMemcpyAsync(d8, src8, cb * dstride * 2, cudaMemcpyHostToDevice, c->copy_stream));
                else
                    CK(c, cudaMemcpy2DAsync(d8, dstride * 2, src8, stride * 2, spb * 2, cb, cudaMemcpyHostToDevice,
                                            c->copy_stream));
            } else if (spb) {
                if (stride == dstride) {
                    CK(c, cudaMemcpyAsync(d_iq + 2 * b0 * dstride, src, cb * dstride * 4, cudaMemcpyHostToDevice, c->copy_stream));
                } else {
                    CK(c, cudaMemcpy2DAsync(d_iq + 2 * b0 * dstride, dstride * 4, src, stride * 4, spb * 4, cb,
                                            cudaMemcpyHostToDevice, c->copy_stream));
                }
            }
            CK(c, cudaEventRecord(c->chunk_events[k], c->copy_stream));
            CK(c, cudaStreamWaitEvent(c->stream, c->chunk_events[k], 0));
            if (spb && iq8) {
                const size_t words = cb * dstride / 2;     // 4 bytes = two samples per thread; dstride % 4 == 0
                cu8_expand_kernel<<<(unsigned)((words + 255) / 256), 256, 0, c->stream>>>(
                    reinterpret_cast<const uint32_t *>(c->d_stage8 + 2 * b0 * dstride),
                    reinterpret_cast<uint2 *>(d_iq + 2 * b0 * dstride), words);
                CK(c, cudaGetLastError());
                c->timing.other_launches++;
            }
            rc = launch_scan(c, (uint32_t)b0, (uint32_t)cb);
            if (rc) { c->cur.active = false; return rc; }
        }
        size_t n = 0;   // (if the pool overflowed the data is resident by now: plain re-run inside)
        rc = scan_resolve(c, true, c->d_frames, room, &n, per_buffer_counts ? c->d_counts : nullptr);
        if (rc && rc != B200ADSB_ERR_CAPACITY)
            return rc;
        if (rc == B200ADSB_ERR_CAPACITY)
            status = rc;
        const size_t ncopy = std::min(n, room);
        if (ncopy)
            CK(c, cudaMemcpyAsync(out + total_frames, c->d_frames, ncopy * sizeof(b200adsb_frame), cudaMemcpyDeviceToHost, c->stream));
        if (per_buffer_counts && nb)
            CK(c, cudaMemcpyAsync(per_buffer_counts + done, c->d_counts, nb * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        // frames carry the buffer index inside this pass: rebase
        if (done)
            for (size_t i = 0; i < ncopy; i++)
                out[total_frames + i].buffer += (uint32_t)done;
        total_frames += n;
        c->next_ordinal += nb;
        done += nb;
    } while (done < n_buffers);
    if (n_out)
        *n_out = total_frames;
    return status;
}
}  // namespace
extern "C" {

int b200adsb_demod_iq_batch(b200adsb_ctx *c, const int16_t *iq, size_t n_buffers, size_t spb,
                            size_t stride, const uint32_t *lengths, b200adsb_frame *out, size_t cap,
                            size_t *n_out, uint32_t *per_buffer_counts)
{
    return host_batch(c, iq, nullptr, n_buffers, spb, stride, lengths, out, cap, n_out, per_buffer_counts);
}

int b200adsb_demod_cu8_batch(b200adsb_ctx *c, const uint8_t *iq_u8, size_t n_buffers, size_t spb,
                             size_t stride, const uint32_t *lengths, b200adsb_frame *out, size_t cap,
                             size_t *n_out, uint32_t *per_buffer_counts)
{
    if (!iq_u8 && n_buffers && spb)
        return B200ADSB_ERR_BAD_ARG;
    return host_batch(c, nullptr, iq_u8, n_buffers, spb, stride, lengths, out, cap, n_out, per_buffer_counts);
}

int16_t b200adsb_cu8_to_cs16(uint8_t v)
{
    return (int16_t)((((float)v - 127.4f) * (1.0f / 128.0f)) * 32767.0f);
}

// ------------------------------------------------------------------ receive loop (double buffered)
int b200adsb_demod_iq_batch_submit(b200adsb_ctx *c, int slot, const int16_t *iq, size_t n_buffers, size_t spb,
                                   size_t stride, const uint32_t *lengths, b200adsb_frame *out, size_t cap,
                                   uint32_t *result)
{
    if (!c || slot < 0 || slot > 1 || (!iq && n_buffers && spb) || (!out && cap) || !result)
        return B200ADSB_ERR_BAD_ARG;
    if (spb > (size_t)kMaxSamples || (stride < spb && n_buffers > 1) || n_buffers == 0)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    b200adsb_ctx::Slot &sl = c->slots[slot];
    if (sl.in_flight)
        return B200ADSB_ERR_STATE;       // wait for the previous batch of this slot first
    const size_t dstride = (spb + 3) & ~(size_t)3;
    rc = grow(c, (unsigned char **)&sl.d_iq, &sl.iq_bytes, std::max<size_t>(n_buffers * dstride * 4, 16), 1);
    if (rc) return rc;
    rc = grow(c, &sl.d_frames, &sl.frames_cap, std::max<size_t>(cap, 1));
    if (rc) return rc;
    if (!sl.d_result)
        CK(c, cudaMalloc((void **)&sl.d_result, 16));
    if (!sl.done)
        CK(c, cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    int16_t *d_iq = reinterpret_cast<int16_t *>(sl.d_iq);
    const uint32_t *d_len = nullptr;
    if (lengths) {
        rc = grow(c, &sl.d_lengths, &sl.lengths_cap, n_buffers);
        if (rc) return rc;
        CK(c, cudaMemcpyAsync(sl.d_lengths, lengths, n_buffers * 4, cudaMemcpyHostToDevice, c->stream));
        d_len = sl.d_lengths;
    }
    if (spb) {
        if (stride == dstride)
            CK(c, cudaMemcpyAsync(d_iq, iq, n_buffers * dstride * 4, cudaMemcpyHostToDevice, c->stream));
        else
            CK(c, cudaMemcpy2DAsync(d_iq, dstride * 4, iq, stride * 4, spb * 4, n_buffers, cudaMemcpyHostToDevice,
                                    c->stream));
    }
    rc = b200adsb_demod_iq_batch_dev_async(c, d_iq, n_buffers, spb, dstride, d_len, sl.d_frames, cap, sl.d_result);
    if (rc) return rc;
    CK(c, cudaMemcpyAsync(result, sl.d_result, 16, cudaMemcpyDeviceToHost, c->stream));
    if (cap)   // the frame count is only known on the device: the whole (small) array travels
        CK(c, cudaMemcpyAsync(out, sl.d_frames, cap * sizeof(b200adsb_frame), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaEventRecord(sl.done, c->stream));
    sl.in_flight = true;
    return B200ADSB_OK;
}

int b200adsb_demod_iq_batch_wait(b200adsb_ctx *c, int slot)
{
    if (!c || slot < 0 || slot > 1)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    b200adsb_ctx::Slot &sl = c->slots[slot];
    if (!sl.in_flight)
        return B200ADSB_ERR_STATE;
    CK(c, cudaEventSynchronize(sl.done));
    sl.in_flight = false;
    return B200ADSB_OK;
}

int b200adsb_demod_iq(b200adsb_ctx *c, const int16_t *iq, size_t n, b200adsb_frame *out, size_t cap,
                      size_t *n_out)
{
    return b200adsb_demod_iq_batch(c, iq, 1, n, n, nullptr, out, cap, n_out, nullptr);
}

int b200adsb_demodulate2400(b200adsb_ctx *c, const uint16_t *data, size_t length, b200adsb_frame *out,
                            size_t cap, size_t *n_out)
{
    if (!c || !data || (!out && cap))
        return B200ADSB_ERR_BAD_ARG;
    if (length > (size_t)kMaxSamples)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    rc = ensure_stage(c, (size_t)kMagLen * 2 + 64);
    if (rc) return rc;
    rc = grow(c, &c->d_frames, &c->frames_cap, std::max<size_t>(cap, 1));
    if (rc) return rc;
    CK(c, cudaMemcpyAsync(c->d_stage, data, (size_t)kMagLen * 2, cudaMemcpyHostToDevice, c->stream));
    rc = scan_begin(c, c->d_stage, true, 1, length, kMagLen, nullptr, c->next_ordinal, 1);
    if (rc) return rc;
    c->next_ordinal += 1;
    size_t n = 0;
    rc = scan_resolve(c, false, c->d_frames, cap, &n, nullptr);
    if (rc && rc != B200ADSB_ERR_CAPACITY)
        return rc;
    const size_t ncopy = std::min(n, cap);
    if (ncopy) {
        CK(c, cudaMemcpyAsync(out, c->d_frames, ncopy * sizeof(b200adsb_frame), cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
    }
    if (n_out)
        *n_out = n;
    return rc;
}

// ------------------------------------------------------------------ icao_filter.rs
uint32_t b200adsb_icao_hash(uint32_t a32)   // src/icao_filter.rs:19-43 (pure function)
{
    uint64_t a = a32, h = 0;
    for (int k = 0; k < 3; k++) {
        h += (a >> (8 * k)) & 0xff;
        h += h << 10;
        h ^= h >> 6;
    }
    h += h << 3;
    h ^= h >> 11;
    h += h << 15;
    return (uint32_t)h & (B200ADSB_ICAO_FILTER_SIZE - 1);
}

int b200adsb_icao_flush(b200adsb_ctx *c)
{
    if (!c)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    CK(c, cudaMemsetAsync(c->d_members, 0, kMemberSlots * 4, c->stream));
    CK(c, cudaMemsetAsync(c->d_counters + C_MEMBERS, 0, 4, c->stream));
    return B200ADSB_OK;   // stream ordered: later calls on this context see the empty filter
}

int b200adsb_icao_filter_add(b200adsb_ctx *c, uint32_t addr)
{
    if (!c)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    filter_add_kernel<<<1, 32, 0, c->stream>>>(c->d_members, c->d_counters, addr);
    CK(c, cudaGetLastError());
    CK(c, cudaStreamSynchronize(c->stream));
    return B200ADSB_OK;
}

int b200adsb_icao_filter_test(b200adsb_ctx *c, uint32_t addr)
{
    if (!c)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    filter_test_kernel<<<1, 32, 0, c->stream>>>(c->d_members, addr, c->d_scalar);
    CK(c, cudaGetLastError());
    uint32_t v = 0;
    CK(c, cudaMemcpyAsync(&v, c->d_scalar, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return (int)v;
}

int b200adsb_icao_snapshot(b200adsb_ctx *c, uint32_t *keys, size_t cap, size_t *n)
{
    if (!c || (!keys && cap))
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    std::vector<uint32_t> tab(kMemberSlots);
    CK(c, cudaMemcpyAsync(tab.data(), c->d_members, kMemberSlots * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    size_t k = 0;
    for (uint32_t v : tab)
        if (v) {
            if (k < cap)
                keys[k] = v;
            k++;
        }
    if (n)
        *n = k;
    return k > cap ? B200ADSB_ERR_CAPACITY : B200ADSB_OK;
}

int b200adsb_icao_restore(b200adsb_ctx *c, const uint32_t *keys, size_t n)
{
    if (!c || (!keys && n) || n > (size_t)B200ADSB_ICAO_FILTER_SIZE)
        return B200ADSB_ERR_BAD_ARG;
    int rc = b200adsb_icao_flush(c);
    if (rc) return rc;
    if (!n)
        return B200ADSB_OK;
    CK(c, cudaMemcpyAsync(c->d_new_keys, keys, n * 4, cudaMemcpyHostToDevice, c->stream));
    filter_restore_kernel<<<1, 1024, 0, c->stream>>>(c->d_members, c->d_counters, c->d_new_keys, (uint32_t)n);
    CK(c, cudaGetLastError());
    CK(c, cudaStreamSynchronize(c->stream));
    return B200ADSB_OK;
}

// ------------------------------------------------------------------ crc.rs / mode_s
int b200adsb_modes_checksum(b200adsb_ctx *c, const uint8_t *msgs, size_t n, size_t bits, uint32_t *out)
{
    if (!c || (!msgs && n) || (!out && n) || (bits != 56 && bits != 112))
        return B200ADSB_ERR_BAD_ARG;   // reference asserts n >= 3 bytes (crc.rs:267)
    if (!n)
        return B200ADSB_OK;
    int rc = bind(c);
    if (rc) return rc;
    rc = ensure_stage(c, n * 14 + n * 4 + 64);
    if (rc) return rc;
    uint8_t *d_m = reinterpret_cast<uint8_t *>(c->d_stage);
    uint32_t *d_o = reinterpret_cast<uint32_t *>(d_m + ((n * 14 + 15) & ~(size_t)15));
    CK(c, cudaMemcpyAsync(d_m, msgs, n * 14, cudaMemcpyHostToDevice, c->stream));
    checksum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_m, (int)n, (int)(bits / 8), c->d_crc256, d_o);
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(out, d_o, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return B200ADSB_OK;
}

int b200adsb_score_modes_messages(b200adsb_ctx *c, const uint8_t *msgs, size_t n, uint8_t *lens,
                                  int32_t *scores)
{
    if (!c || (!msgs && n) || (!lens && n) || (!scores && n))
        return B200ADSB_ERR_BAD_ARG;
    if (!n)
        return B200ADSB_OK;
    if (c->cur.active)
        return B200ADSB_ERR_STATE;
    int rc = bind(c);
    if (rc) return rc;
    const int per_tile = 1024;
    const size_t n_tiles = (n + per_tile - 1) / per_tile;
    rc = ensure_stage(c, n * 14 + 64);
    if (rc) return rc;
    rc = ensure_pool(c, std::max<size_t>(n, 4096));
    if (rc) return rc;
    rc = ensure_tiles(c, n_tiles);
    if (rc) return rc;
    rc = grow(c, &c->d_frames, &c->frames_cap, 1);
    if (rc) return rc;
    uint8_t *d_m = reinterpret_cast<uint8_t *>(c->d_stage);
    CK(c, cudaMemcpyAsync(d_m, msgs, n * 14, cudaMemcpyHostToDevice, c->stream));
    rc = reset_scan_counters(c);
    if (rc) return rc;
    classify_msgs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        d_m, (int)n, c->d_crc256, c->d_rec, c->d_tile_dir, per_tile, c->d_counters, c->d_ev_keys,
        c->d_ev_ord, c->d_ev_used, kEvSlots - 1, c->next_ordinal);
    CK(c, cudaGetLastError());
    rc = read_counters(c);
    if (rc) return rc;
    if (c->h_counters[C_FLAGS] & F_EV_OVF) {
        clear_events(c);
        return B200ADSB_ERR_EVENTS;
    }
    Pending &q = c->cur;
    q = Pending();
    q.active = true;
    q.in = d_m;
    q.n_buffers = (uint32_t)n_tiles;
    q.spb = per_tile;
    q.stride = per_tile;
    q.T = per_tile;
    q.tpb = 1;
    q.n_tiles = (uint32_t)n_tiles;
    q.ord_first = c->next_ordinal;
    q.ord_stride = 1;
    q.msgs = d_m;
    c->next_ordinal += n_tiles;
    size_t nf = 0;
    rc = resolve_run(c, c->d_frames, 0, &nf, nullptr);
    if (rc && rc != B200ADSB_ERR_CAPACITY)
        return rc;
    CK(c, cudaMemcpyAsync(scores, c->d_rec_score, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; i++) {
        // MsgLen from the DF bit (mode_s/mod.rs:42-46); None for an all-zero message (:51-53)
        const uint8_t *m = msgs + 14 * i;
        bool any = false;
        for (int k = 0; k < 14; k++)
            any |= m[k] != 0;
        lens[i] = !any ? 0 : ((m[0] & 0x80) ? 14 : 7);
    }
    return B200ADSB_OK;
}

/* single-message forms with the reference's own shapes (crc.rs:263, mode_s/mod.rs:14,34) */
int b200adsb_modes_checksum_one(b200adsb_ctx *c, const uint8_t *msg, size_t n_bytes, size_t bits, uint32_t *out)
{
    // the reference asserts bits % 8 == 0 and bits / 8 <= msg.len() with at least 3 bytes (crc.rs:264-267)
    if (!c || !msg || !out || (bits != 56 && bits != 112) || n_bytes < bits / 8)
        return B200ADSB_ERR_BAD_ARG;
    uint8_t m14[14] = {0};
    memcpy(m14, msg, std::min<size_t>(n_bytes, 14));
    return b200adsb_modes_checksum(c, m14, 1, bits, out);
}

int b200adsb_score_modes_message(b200adsb_ctx *c, const uint8_t *msg, size_t n_bytes, int *msglen, int *score)
{
    if (!c || !msg || !msglen || !score)
        return B200ADSB_ERR_BAD_ARG;
    *msglen = 0;
    *score = 0;
    // mode_s/mod.rs:35-49: None when the slice is shorter than the length its DF announces
    if (n_bytes < (size_t)B200ADSB_MODES_SHORT_MSG_BYTES)
        return B200ADSB_OK;
    const size_t need = (msg[0] & 0x80) ? B200ADSB_MODES_LONG_MSG_BYTES : B200ADSB_MODES_SHORT_MSG_BYTES;
    if (n_bytes < need)
        return B200ADSB_OK;
    // :51 looks at every byte of the slice it was given (the scan always hands over 14)
    bool any = false;
    for (size_t k = 0; k < n_bytes; k++)
        any |= msg[k] != 0;
    if (!any)
        return B200ADSB_OK;
    uint8_t m14[14] = {0};
    memcpy(m14, msg, std::min<size_t>(n_bytes, 14));
    uint8_t len = 0;
    int32_t sc = 0;
    const int rc = b200adsb_score_modes_messages(c, m14, 1, &len, &sc);
    if (rc)
        return rc;
    *msglen = (int)need;
    *score = sc;
    return B200ADSB_OK;
}

uint32_t b200adsb_getbits(const uint8_t *data, size_t firstbit_1idx, size_t lastbit_1idx)   /* mode_s/mod.rs:14-30 */
{
    uint32_t ans = 0;
    if (!data || firstbit_1idx == 0)
        return 0;
    for (size_t bit = firstbit_1idx - 1; bit + 1 <= lastbit_1idx; bit++)
        ans = (ans << 1) | ((data[bit / 8] >> (7 - bit % 8)) & 1u);
    return ans;
}

/* acknowledges a failed enqueue-only batch (d_result[1] != 0): the batches queued after this call
 * commit again.  Stream ordered; any synchronous demodulation call does the same. */
int b200adsb_async_acknowledge(b200adsb_ctx *c)
{
    if (!c)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    CK(c, cudaMemsetAsync(c->d_counters + C_STICKY, 0, 4, c->stream));
    return B200ADSB_OK;
}

/* dump1090_rs/src/main.rs:174-176: one "*{hex};\n" line per frame (the AVR text the
 * reference writes to its TCP clients).  Pure host formatting of frames already demodulated. */
int b200adsb_format_avr(const b200adsb_frame *frames, size_t n, char *out, size_t cap, size_t *len)
{
    if ((!frames && n) || (!out && cap))
        return B200ADSB_ERR_BAD_ARG;
    static const char *hexd = "0123456789abcdef";
    size_t need = 0;
    for (size_t i = 0; i < n; i++)
        need += 3 + 2 * (size_t)frames[i].len;
    if (len)
        *len = need;
    if (need > cap)
        return B200ADSB_ERR_CAPACITY;
    size_t o = 0;
    for (size_t i = 0; i < n; i++) {
        out[o++] = '*';
        for (unsigned k = 0; k < frames[i].len; k++) {
            out[o++] = hexd[frames[i].msg[k] >> 4];
            out[o++] = hexd[frames[i].msg[k] & 15];
        }
        out[o++] = ';';
        out[o++] = '\n';
    }
    return B200ADSB_OK;
}

/* test hook: compares the scan kernel's fast magnitude with the IEEE-intrinsic form of
 * utils.rs:47-55 on all 2^32 (re, im) inputs, on the GPU. */
int b200adsb_debug_mag_sweep(b200adsb_ctx *c, uint64_t *mismatches, uint32_t *first_bad)
{
    if (!c || !mismatches || !first_bad)
        return B200ADSB_ERR_BAD_ARG;
    int rc = bind(c);
    if (rc) return rc;
    rc = ensure_stage(c, 64);
    if (rc) return rc;
    unsigned long long *d_m = reinterpret_cast<unsigned long long *>(c->d_stage);
    uint32_t *d_f = reinterpret_cast<uint32_t *>(d_m + 1);
    CK(c, cudaMemsetAsync(d_m, 0, 8, c->stream));
    CK(c, cudaMemsetAsync(d_f, 0xff, 4, c->stream));
    mag_sweep_kernel<<<(1u << 24) / 256, 256, 0, c->stream>>>(d_m, d_f);
    CK(c, cudaGetLastError());
    unsigned long long m = 0;
    CK(c, cudaMemcpyAsync(&m, d_m, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaMemcpyAsync(first_bad, d_f, 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    *mismatches = m;
    return B200ADSB_OK;
}

/* test hook: the stage-1 records of the pending batch (valid between b200adsb_scan_batch_dev and
 * b200adsb_resolve_batch_dev): for every position that passed the preamble gates, in (buffer, j)
 * order, the batch buffer index and the six record words {j, w[5]} (w[t-4] = kind<<29 | key of
 * try-phase t) -- the per-stage parity point of SURVEY.md section 7 steps 4-5 (survivor set and
 * per-(j, t) classification, src/demod_2400.rs:127-146,158-189, src/mode_s/mod.rs:34-139). */
int b200adsb_debug_records(b200adsb_ctx *c, uint32_t *buffers, uint32_t *rec6, size_t cap, size_t *n_out)
{
    if (!c || !n_out || (cap && (!buffers || !rec6)))
        return B200ADSB_ERR_BAD_ARG;
    if (!c->cur.active)
        return B200ADSB_ERR_STATE;
    int rc = bind(c);
    if (rc) return rc;
    const Pending &q = c->cur;
    rc = read_counters(c);
    if (rc) return rc;
    const size_t used = c->h_counters[C_POOL];
    std::vector<uint2> dir(q.n_tiles);
    std::vector<uint32_t> pool(6 * std::max<size_t>(used, 1));
    if (q.n_tiles)
        CK(c, cudaMemcpy(dir.data(), c->d_tile_dir, q.n_tiles * sizeof(uint2), cudaMemcpyDeviceToHost));
    if (used)
        CK(c, cudaMemcpy(pool.data(), c->d_rec, used * 24, cudaMemcpyDeviceToHost));
    size_t n = 0;
    for (uint32_t t = 0; t < q.n_tiles; t++)
        for (uint32_t i = 0; i < dir[t].y; i++, n++) {
            if (n >= cap)
                continue;
            buffers[n] = t / (uint32_t)std::max(q.tpb, 1);
            memcpy(rec6 + 6 * n, pool.data() + 6 * ((size_t)dir[t].x + i), 24);
        }
    *n_out = n;
    return n > cap ? B200ADSB_ERR_CAPACITY : B200ADSB_OK;
}

/* test hook (pure host): the CRC-24 field tables the scan kernel uses, so that CPU
 * tests can check them against crc.rs semantics without a GPU.  out: 840 + 256 u32 */
int b200adsb_debug_crc_tabs(uint32_t *out)
{
    if (!out)
        return B200ADSB_ERR_BAD_ARG;
    build_crc_tabs(out, out + kTabWords);
    return kTabWords;
}

/* test hook (pure host): the per-lane shuffle form of the same tables, 7 x 32 u32 (kernels.cuh::a112_sh) */
int b200adsb_debug_crc_lane_tabs(uint32_t *out)
{
    if (!out)
        return B200ADSB_ERR_BAD_ARG;
    uint32_t t[kTabWords], t256[256];
    build_crc_tabs(t, t256, out);
    return kLaneTabs * 32;
}

}  // extern "C"
