/*
 * include/b200adsb.h -- C ABI of libb200adsb.so, the B200-native (sm_100a CUDA)
 * replacement of the dump1090_rs demodulation hot path
 *
 *     CS16 IQ -> magnitude -> 8 us preamble gate -> 5-phase PPM bit slicing
 *             -> Mode-S CRC-24 / scoring -> 56/112-bit frames
 *
 * Each entry point names the reference interface it replaces (file:line in
 * rsadsb/dump1090_rs @ 94a0e4d).  Plain pointers and sizes only; no C++ or
 * torch types cross this boundary.  Every function returns a status code
 * (B200ADSB_OK == 0, negative on error) and never aborts or unwinds.
 *
 * Threading: one context == one ordered stream of buffers with one ICAO
 * address filter (the reference's process-wide tables, src/icao_filter.rs:8-9).
 * Calls on one context must be serialised by the caller; contexts are
 * independent.  Results are ordered by (buffer, j) exactly like the reference's
 * Vec<ModeSMessage> (src/demod_2400.rs:121,207).
 *
 * There is no CPU fallback: every call that computes runs CUDA kernels on the
 * context's device and fails with B200ADSB_ERR_CUDA if that is impossible.
 */
#ifndef B200ADSB_H
#define B200ADSB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/lib.rs:22-26 */
#define B200ADSB_MODES_MAG_BUF_SAMPLES 131072
#define B200ADSB_TRAILING_SAMPLES 326
#define B200ADSB_MAG_DATA_LEN (B200ADSB_TRAILING_SAMPLES + B200ADSB_MODES_MAG_BUF_SAMPLES)
#define B200ADSB_MODES_LONG_MSG_BYTES 14
#define B200ADSB_MODES_SHORT_MSG_BYTES 7
/* src/icao_filter.rs:5-6 */
#define B200ADSB_ICAO_FILTER_SIZE 4096
#define B200ADSB_ICAO_FILTER_ADSB_NT (1u << 25)

enum {
    B200ADSB_OK = 0,
    B200ADSB_ERR_BAD_ARG = -1,   /* null pointer, n > 131072 (reference panics, lib.rs:48), ... */
    B200ADSB_ERR_CAPACITY = -2,  /* caller's output array too small; *n_out holds the need   */
    B200ADSB_ERR_CUDA = -3,      /* CUDA runtime error; see b200adsb_last_error()             */
    B200ADSB_ERR_NOMEM = -4,
    B200ADSB_ERR_STATE = -5,     /* calls out of order (resolve before scan, ...)             */
    B200ADSB_ERR_EVENTS = -6     /* more distinct new addresses in one batch than event slots */
};

/* context options (b200adsb_ctx_set_option) */
enum {
    B200ADSB_OPT_TILE = 1,        /* output positions per thread block tile (multiple of 8, <= 8184) */
    B200ADSB_OPT_POOL_SHIFT = 2,  /* candidate pool = positions >> shift (grown on demand)    */
    B200ADSB_OPT_PROFILE = 3,     /* 1: bracket kernels with CUDA events (b200adsb_timing)    */
    B200ADSB_OPT_H2D_CHUNK = 4,   /* buffers per host->device pipeline chunk (host batch API) */
    B200ADSB_OPT_CARRY = 5        /* 1: stream continuity -- the 326 leading MagnitudeBuffer slots of
                                     a buffer carry the previous buffer's last samples (what the C
                                     dump1090 did and the reference drops, lib.rs:24,47-50,
                                     utils.rs:44); recovers frames that straddle two buffers.  Off by
                                     default: it changes the output relative to the reference.     */
};

typedef struct b200adsb_ctx b200adsb_ctx;

/* One decoded frame: what ModeSMessage exposes through buffer()
 * (src/demod_2400.rs:93-112) plus where it was found.  28 bytes. */
typedef struct {
    uint8_t msg[B200ADSB_MODES_LONG_MSG_BYTES]; /* msg[..len] == ModeSMessage::buffer() */
    uint8_t len;        /* 7 (MsgLen::Short) or 14 (MsgLen::Long)                        */
    uint8_t phase;      /* winning try_phase, 4..8 (private in the reference)            */
    int16_t score;      /* score_modes_message value of the winner (private there)       */
    uint16_t reserved;
    uint32_t j;         /* index into MagnitudeBuffer.data where the preamble starts     */
    uint32_t buffer;    /* index of the IQ buffer inside the batch                       */
} b200adsb_frame;

/* accumulated device time per stage since the last reset (OPT_PROFILE=1) */
typedef struct {
    double scan_ms;     /* fused magnitude+preamble+slice+CRC kernel                     */
    double resolve_ms;  /* event finalise + filter resolve + ordered emit                */
    double h2d_ms, d2h_ms;
    uint64_t scan_launches, other_launches;
    uint64_t samples;   /* IQ samples pushed through the scan kernel                     */
    uint64_t candidates;/* positions that passed all preamble gates                      */
} b200adsb_timing;

/* ------------------------------------------------------------------ life cycle */
int b200adsb_version(void);
const char *b200adsb_strerror(int status);
const char *b200adsb_last_error(const b200adsb_ctx *ctx); /* last CUDA error text */

/* device: CUDA ordinal.  stream: a cudaStream_t to launch on (e.g. torch's
 * current stream) or NULL for a private non-blocking stream. */
int b200adsb_ctx_create(b200adsb_ctx **out, int device, void *stream);
void b200adsb_ctx_destroy(b200adsb_ctx *ctx);
int b200adsb_ctx_set_option(b200adsb_ctx *ctx, int option, int64_t value);
int b200adsb_ctx_sync(b200adsb_ctx *ctx);
int b200adsb_timing_get(b200adsb_ctx *ctx, b200adsb_timing *out, int reset);

/* pinned host memory for the host-buffer entry points (plain malloc'd memory
 * works too, only slower) */
void *b200adsb_host_alloc(size_t bytes);
void b200adsb_host_free(void *p);

/* ------------------------------------------------ reference surface, one buffer */

/* utils::to_mag(&[Complex<i16>]) -> MagnitudeBuffer        (src/utils.rs:43-58)
 * iq_re_im: n interleaved (re, im) int16 pairs in memory order (host memory).
 * data: host array of B200ADSB_MAG_DATA_LEN u16 (MagnitudeBuffer.data,
 * src/lib.rs:31); data[0..326) = 0, data[326+k] = magnitude of sample k,
 * the rest 0 (MagnitudeBuffer::default, lib.rs:36-44).  *length = n. */
int b200adsb_to_mag(b200adsb_ctx *ctx, const int16_t *iq_re_im, size_t n, uint16_t *data,
                    size_t *length);

/* demod_2400::demodulate2400(&MagnitudeBuffer) -> Vec<ModeSMessage>
 *                                                          (src/demod_2400.rs:115-212)
 * data/length: a MagnitudeBuffer (host).  Uses and updates the context filter. */
int b200adsb_demodulate2400(b200adsb_ctx *ctx, const uint16_t *data, size_t length,
                            b200adsb_frame *out, size_t cap, size_t *n_out);

/* to_mag + demodulate2400 fused (the pair every caller issues: main.rs:166-167,
 * tests/test.rs:11-13, benches/demod_benchmark.rs:10-11); the magnitude never
 * leaves the chip. */
int b200adsb_demod_iq(b200adsb_ctx *ctx, const int16_t *iq_re_im, size_t n,
                      b200adsb_frame *out, size_t cap, size_t *n_out);

/* --------------------------------------------------------------- batched stream */

/* n_buffers consecutive buffers of one stream (buffer b starts at
 * iq + 2*b*stride_samples int16, has samples_per_buffer samples, or lengths[b]
 * if lengths != NULL), processed as the reference would process them one after
 * the other with one filter.  Host pointers; H2D/D2H are pipelined inside.
 * per_buffer_counts (nullable, host): frames found in each buffer. */
int b200adsb_demod_iq_batch(b200adsb_ctx *ctx, const int16_t *iq_re_im, size_t n_buffers,
                            size_t samples_per_buffer, size_t stride_samples,
                            const uint32_t *lengths, b200adsb_frame *out, size_t cap,
                            size_t *n_out, uint32_t *per_buffer_counts);

/* Opt-in ingest for 8-bit front ends (RTL-SDR, the reference's default --driver): the same call on unsigned
 * 8-bit (I, Q) pairs as the radio delivers them.  They cross PCIe as 2 bytes per sample and are expanded on
 * the device with the conversion SoapySDR's RTL-SDR module applies when the reference asks it for CS16
 * (main.rs:143): int16((float(u8) - 127.4f) * (1.0f / 128.0f) * 32767.0f) -- b200adsb_cu8_to_cs16 is that
 * function on the host -- so the frames are those of b200adsb_demod_iq_batch on the converted samples. */
int b200adsb_demod_cu8_batch(b200adsb_ctx *ctx, const uint8_t *iq_u8, size_t n_buffers,
                             size_t samples_per_buffer, size_t stride_samples, const uint32_t *lengths,
                             b200adsb_frame *out, size_t cap, size_t *n_out, uint32_t *per_buffer_counts);
int16_t b200adsb_cu8_to_cs16(uint8_t v);

/* Same with the IQ already resident in device memory (d_iq) and frames left in
 * device memory (d_out, cap entries); *n_out is read back.  d_lengths nullable
 * (device).  d_per_buffer_counts nullable (device, n_buffers u32). */
int b200adsb_demod_iq_batch_dev(b200adsb_ctx *ctx, const int16_t *d_iq, size_t n_buffers,
                                size_t samples_per_buffer, size_t stride_samples,
                                const uint32_t *d_lengths, b200adsb_frame *d_out, size_t cap,
                                size_t *n_out, uint32_t *d_per_buffer_counts);

/* Enqueue-only form of b200adsb_demod_iq_batch_dev for device-resident pipelines (the
 * reference's loop `stream.read -> to_mag -> demodulate2400`, main.rs:161-167, with the
 * consumer of the frames on the device or reading them later): the call returns as soon as
 * the batch is queued on the context's stream; nothing is read back.  d_result (device,
 * 4 x uint32): [0] frames written to d_out in (buffer, j) order, [1] 0, or non-zero when
 * the batch failed -- bit 0: the optimistic candidate pool overflowed, bit 1: the event table /
 * exchange buffer overflowed, bit 4: another rank of a sharded stream failed its batch, bit 3: an
 * EARLIER queued batch failed and has not been acknowledged.  A failed batch is NOT committed to
 * the filter and wrote no valid frames, and neither is any batch queued behind it (their d_result[1]
 * has bit 3 set), so the filter is exactly what it was before the first failed batch: rerun that batch
 * and the ones queued after it with the synchronous call (which grows the pool and acknowledges the
 * failure; b200adsb_async_acknowledge does only the latter); [2] positions that passed the preamble
 * gates; [3] 1 when more than `cap` frames were found.  Batches on one context execute in call
 * order with one filter, exactly like the synchronous form. */
int b200adsb_demod_iq_batch_dev_async(b200adsb_ctx *ctx, const int16_t *d_iq, size_t n_buffers,
                                      size_t samples_per_buffer, size_t stride_samples,
                                      const uint32_t *d_lengths, b200adsb_frame *d_out, size_t cap,
                                      uint32_t *d_result);

int b200adsb_async_acknowledge(b200adsb_ctx *ctx);

/* The receive loop, double buffered (main.rs:161-167: `stream.read` -> `to_mag` -> `demodulate2400`, where
 * the reference's demodulation blocks the next read): submit queues H2D of the host batch (pinned memory
 * for true asynchrony), the enqueue-only demodulation and the D2H of the 4-word outcome (as d_result above)
 * and of `cap` frame slots into `out` / `result` (host, pinned), and returns; the caller reads the next
 * batch from its source while the GPU works, then waits for the slot (0 or 1) and consumes result/out.
 * A failed batch (result[1] != 0) is redone with b200adsb_demod_iq_batch -- and so is a batch submitted
 * behind it (its result[1] has bit 3 set), in order. */
int b200adsb_demod_iq_batch_submit(b200adsb_ctx *ctx, int slot, const int16_t *iq_re_im, size_t n_buffers,
                                   size_t samples_per_buffer, size_t stride_samples, const uint32_t *lengths,
                                   b200adsb_frame *out, size_t cap, uint32_t *result);
int b200adsb_demod_iq_batch_wait(b200adsb_ctx *ctx, int slot);

/* ---- split form for a stream sharded over several GPUs (one context per rank).
 * scan:    stage 1 on this rank's buffers; local buffer b has stream ordinal
 *          first_ordinal + b*ordinal_stride (round robin: first=rank, stride=world).
 * events:  the ICAO add-events (key, first ordinal) this rank saw -- a few per
 *          buffer -- to be all-gathered between ranks and imported everywhere.
 * resolve: stage 2, identical filter evolution on every rank.               */
int b200adsb_scan_batch_dev(b200adsb_ctx *ctx, const int16_t *d_iq, size_t n_buffers,
                            size_t samples_per_buffer, size_t stride_samples,
                            const uint32_t *d_lengths, uint64_t first_ordinal,
                            uint64_t ordinal_stride);
int b200adsb_events_count(b200adsb_ctx *ctx, size_t *n);
/* pairs: 2*cap u64 in DEVICE memory: (key, ordinal) per event */
int b200adsb_events_export_dev(b200adsb_ctx *ctx, uint64_t *d_pairs, size_t cap, size_t *n);
int b200adsb_events_import_dev(b200adsb_ctx *ctx, const uint64_t *d_pairs, size_t n);
/* the same exchange without host round trips: rows[0] = (count, bad-batch flags of this rank),
 * rows[1..] = (key, ordinal); all-gather the fixed-size row blocks, then merge every block but
 * this rank's own.  A rank with more than rows-1 events, or whose scan overflowed, fails the batch
 * on EVERY rank (itself included): the later resolve returns B200ADSB_ERR_EVENTS / reports it in
 * d_result[1], and no rank commits -- the ranks' filters stay identical. */
int b200adsb_events_pack_dev(b200adsb_ctx *ctx, uint64_t *d_rows, size_t rows_cap);
int b200adsb_events_import_packed_dev(b200adsb_ctx *ctx, const uint64_t *d_gathered, size_t n_ranks,
                                      size_t rows_per_rank, size_t skip_rank);
/* The same exchange fused with its transport (no collective library call on the scan -> resolve path):
 * every rank owns a buffer of b200adsb_events_symm_words(n_ranks, rows_per_rank) u64 in symmetric
 * (peer-mapped, zero-initialised) device memory; d_peer_bufs is a DEVICE array of the n_ranks base
 * pointers as seen from this rank.  push packs this rank's events and stores them into its slot of every
 * rank's buffer over NVLink, then raises its flag there; import waits (on the device, bounded) for all
 * n_ranks flags of this epoch and merges.  epoch: 1, 2, 3, ... identical on all ranks, one per exchange.
 * force_flags: bad-batch flags to publish even though the local counters are clean (a rank whose scan
 * failed before the exchange).  A peer that never arrives fails the batch (bit 4) instead of hanging. */
size_t b200adsb_events_symm_words(size_t n_ranks, size_t rows_per_rank);
int b200adsb_events_push_symm_dev(b200adsb_ctx *ctx, uint64_t *const *d_peer_bufs, size_t rank, size_t n_ranks,
                                  size_t rows_per_rank, uint64_t epoch, uint32_t force_flags);
int b200adsb_events_import_symm_dev(b200adsb_ctx *ctx, uint64_t *d_local_buf, size_t rank, size_t n_ranks,
                                    size_t rows_per_rank, uint64_t epoch);
int b200adsb_resolve_batch_dev(b200adsb_ctx *ctx, b200adsb_frame *d_out, size_t cap,
                               size_t *n_out, uint32_t *d_per_buffer_counts);

/* frames: the reference emits ONE stream in (buffer, j) order (dump1090_rs/src/main.rs:166-200).  Each
 * rank packs its frames (local buffer indices) into a fixed-size block of 1 + rows_cap rows of
 * sizeof(b200adsb_frame): row 0 = {u32 count, 0...}; the blocks are all-gathered; merge writes the single
 * ordered stream with GLOBAL buffer indices (round robin: global = local * n_ranks + rank) to d_out and
 * d_n_out[0] = frames, d_n_out[1] = 1 if a rank had more than rows_cap frames or the total exceeds cap.
 * count: *d_count (device) if d_count != NULL, else `count`.  Enqueue-only, on `stream` (a cudaStream_t;
 * NULL = the context's stream): the gather depends on nothing but the batch's resolve and nothing but the
 * emitter waits for it, so it can run on a side stream while the next batch is scanned. */
int b200adsb_frames_pack_dev(b200adsb_ctx *ctx, void *stream, const b200adsb_frame *d_frames,
                             const uint32_t *d_count, size_t count, b200adsb_frame *d_block, size_t rows_cap);
int b200adsb_frames_merge_dev(b200adsb_ctx *ctx, void *stream, const b200adsb_frame *d_gathered, size_t n_ranks,
                              size_t rows_cap, b200adsb_frame *d_out, size_t cap, uint32_t *d_n_out);

/* The gather fused with its transport (as b200adsb_events_push/import_symm_dev): every rank owns
 * b200adsb_frames_symm_bytes(n_ranks, rows_cap) bytes of zero-initialised symmetric memory; push stores this
 * rank's block into every rank's buffer over NVLink and raises its flag, merge waits (on the device,
 * bounded) for all flags of the epoch (1, 2, 3, ... one per gather, identical on all ranks) and writes the
 * ordered stream; d_n_out[1] = 3 when a peer never arrived.  d_ticket: one zeroed device word per caller.
 * A rank must not start the event exchange of the batch whose gather has epoch e before its own merge of
 * epoch e - 2 has finished (the blocks alternate between two parities); sharded.py orders that with events. */
size_t b200adsb_frames_symm_bytes(size_t n_ranks, size_t rows_cap);
int b200adsb_frames_push_symm_dev(b200adsb_ctx *ctx, void *stream, const b200adsb_frame *d_frames,
                                  const uint32_t *d_count, size_t count, void *const *d_peer_bufs, size_t rank,
                                  size_t n_ranks, size_t rows_cap, uint64_t epoch, uint32_t *d_ticket);
int b200adsb_frames_merge_symm_dev(b200adsb_ctx *ctx, void *stream, void *d_local_buf, size_t n_ranks,
                                   size_t rows_cap, uint64_t epoch, b200adsb_frame *d_out, size_t cap,
                                   uint32_t *d_n_out);

/* enqueue-only forms (no host round trip; outcome in d_result as for
 * b200adsb_demod_iq_batch_dev_async): scan_async -> events_pack -> all-gather ->
 * events_import_packed -> resolve_async, all on the context's stream */
int b200adsb_scan_batch_dev_async(b200adsb_ctx *ctx, const int16_t *d_iq, size_t n_buffers,
                                  size_t samples_per_buffer, size_t stride_samples,
                                  const uint32_t *d_lengths, uint64_t first_ordinal,
                                  uint64_t ordinal_stride);
int b200adsb_resolve_batch_dev_async(b200adsb_ctx *ctx, b200adsb_frame *d_out, size_t cap,
                                     uint32_t *d_result);

/* --------------------------------------------------------------- icao_filter.rs */
int b200adsb_icao_flush(b200adsb_ctx *ctx);                    /* icao_flush       :11-17 */
uint32_t b200adsb_icao_hash(uint32_t a);                       /* icao_hash        :19-43 */
int b200adsb_icao_filter_add(b200adsb_ctx *ctx, uint32_t addr);/* icao_filter_add  :46-62 */
int b200adsb_icao_filter_test(b200adsb_ctx *ctx, uint32_t addr);/* icao_filter_test :65-97; 1/0 */
/* checkpoint / resume of the filter (the reference keeps it in RAM only) */
int b200adsb_icao_snapshot(b200adsb_ctx *ctx, uint32_t *keys, size_t cap, size_t *n);
int b200adsb_icao_restore(b200adsb_ctx *ctx, const uint32_t *keys, size_t n);

/* ------------------------------------------------------- crc.rs / mode_s/mod.rs */
/* modes_checksum(&[u8], bits) (src/crc.rs:263-282) for n messages of 14 bytes
 * each (host), bits = 56 or 112; syndromes to out[n].  Runs the device CRC. */
int b200adsb_modes_checksum(b200adsb_ctx *ctx, const uint8_t *msgs14, size_t n, size_t bits,
                            uint32_t *out);
/* score_modes_message (src/mode_s/mod.rs:34-139) applied to n 14-byte messages
 * in order, against and updating the context filter.  lens[i] = 7/14 (0 for
 * None), scores[i] = score. */
int b200adsb_score_modes_messages(b200adsb_ctx *ctx, const uint8_t *msgs14, size_t n,
                                  uint8_t *lens, int32_t *scores);

/* The same two with the reference's own shapes, one message per call (thin wrappers over the batched
 * forms; the context stands for the process-wide filter):
 *   modes_checksum(message: &[u8], bits: usize) -> u32                 (src/crc.rs:263-282)
 *   score_modes_message(msg: &[u8]) -> Option<(MsgLen, i32)>           (src/mode_s/mod.rs:34-139)
 *       *msglen = 7 / 14, or 0 for None (then *score is 0)
 *   getbits(data, firstbit_1idx, lastbit_1idx)  (pure, host)           (src/mode_s/mod.rs:14-30) */
int b200adsb_modes_checksum_one(b200adsb_ctx *ctx, const uint8_t *msg, size_t n_bytes, size_t bits,
                                uint32_t *out);
int b200adsb_score_modes_message(b200adsb_ctx *ctx, const uint8_t *msg, size_t n_bytes, int *msglen,
                                 int *score);
uint32_t b200adsb_getbits(const uint8_t *data, size_t firstbit_1idx, size_t lastbit_1idx);

/* ------------------------------------------------------- dump1090_rs/src/main.rs:174-176 */
/* "*{hex};\n" per frame, the AVR text the reference binary sends to its TCP clients (host
 * formatting only; the listener itself is out of scope).  *len = bytes needed/written. */
int b200adsb_format_avr(const b200adsb_frame *frames, size_t n, char *out, size_t cap, size_t *len);

/* test hook: exhaustive GPU comparison of the scan kernel's fast magnitude arithmetic
 * with the IEEE-intrinsic statement of src/utils.rs:47-55 over all 2^32 inputs. */
int b200adsb_debug_mag_sweep(b200adsb_ctx *ctx, uint64_t *mismatches, uint32_t *first_bad);

/* test hook: the stage-1 records of the pending batch, valid between b200adsb_scan_batch_dev and
 * b200adsb_resolve_batch_dev: for every position that passed the preamble gates
 * (src/demod_2400.rs:127-146), in (buffer, j) order, its batch buffer index and the six words
 * {j, w[0..5)}, w[t-4] = kind<<29 | key of try-phase t (the stateless part of
 * src/mode_s/mod.rs:34-139; kinds as in oracle/dump1090_oracle.h). */
int b200adsb_debug_records(b200adsb_ctx *ctx, uint32_t *buffers, uint32_t *rec6, size_t cap, size_t *n);

/* test hook, pure host: the 840-word CRC-24 field tables (8/8/6- and 8/3-bit chunks; the scan kernel uses the
 * 5-bit chunking below)
 * followed by the 256-entry byte table (src/crc.rs:3-260); returns 840. */
int b200adsb_debug_crc_tabs(uint32_t *out);
/* the same field sums as 7 tables of 32 entries for warp-shuffle lookup (5-bit chunks); returns 224 */
int b200adsb_debug_crc_lane_tabs(uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif /* B200ADSB_H */
