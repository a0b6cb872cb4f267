/*
 * oracle/dump1090_oracle.h -- CPU restatement of the dump1090_rs hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or as
 * the timed CPU baseline.  The product path (libb200adsb.so) never links,
 * loads or calls it.
 *
 * Parity status: PINNED.  The restatement reproduces all 16 live golden frames
 * of the reference's own integration tests (tests/test.rs:22-28, :35-40,
 * :49-56) on the three committed captures, see tests/test_oracle_golden.py.
 * The reference itself (Rust) cannot be compiled in this image (no rustc /
 * cargo), so there is no oracle/_ref.
 *
 * Every function cites the reference file:line it follows
 * (paths relative to the rsadsb/dump1090_rs tree @ 94a0e4d).
 */
#ifndef DUMP1090_ORACLE_H
#define DUMP1090_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAG_BUF_SAMPLES 131072 /* src/lib.rs:22 */
#define ORC_TRAILING_SAMPLES 326   /* src/lib.rs:24 */
#define ORC_LONG_MSG_BYTES 14      /* src/lib.rs:25 */
#define ORC_SHORT_MSG_BYTES 7      /* src/lib.rs:26 */
#define ORC_ICAO_FILTER_SIZE 4096  /* src/icao_filter.rs:5 */
#define ORC_ICAO_FILTER_ADSB_NT (1u << 25) /* src/icao_filter.rs:6 */

/* src/lib.rs:29-34 */
typedef struct {
    uint16_t data[ORC_TRAILING_SAMPLES + ORC_MAG_BUF_SAMPLES];
    size_t length;
} orc_magbuf;

/* src/icao_filter.rs:8-9 (the two process-wide tables), made an explicit
 * object so that several independent streams can be checked in one process. */
typedef struct {
    uint32_t a[ORC_ICAO_FILTER_SIZE];
    uint32_t b[ORC_ICAO_FILTER_SIZE];
    uint64_t full_events; /* "icao24 hash table full" occurrences */
} orc_filter;

/* src/demod_2400.rs:93-102 plus the (j, phase) the reference keeps private. */
typedef struct {
    uint8_t msg[ORC_LONG_MSG_BYTES];
    uint8_t len;   /* 7 or 14: ModeSMessage::buffer() length */
    uint8_t phase; /* winning try_phase 4..8 */
    int32_t score;
    uint32_t j;    /* index into MagnitudeBuffer.data */
    double signal_level;
} orc_frame;

/* Stateless classification of one (j, try_phase) decode: the order-free
 * restatement of score_modes_message used by the two-pass / multi-rank form
 * (SURVEY.md Appendix A.5/A.6).  kind values are shared with the CUDA path. */
enum {
    ORC_K_NONE = 0,       /* score -2 (or None) whatever the filter holds      */
    ORC_K_PAR_SHORT = 1,  /* DF0/4/5: key=syn56;  member ? 1000 : -1, len 7    */
    ORC_K_DF11_IID0 = 2,  /* DF11 syn56==0: key=addr; member ? 1600 : add,750  */
    ORC_K_DF11_IID = 3,   /* DF11 iid!=0:  key=addr; member ? 1000 : -1        */
    ORC_K_DF17 = 4,       /* DF17 syn112==0: key=addr; member ? 1800 : add,1400 */
    ORC_K_DF18 = 5,       /* DF18 syn112==0: same, but adds addr|ADSB_NT       */
    ORC_K_PAR_LONG = 6    /* DF16/20/21/24..31: key=syn112; member?1000:-2     */
};

typedef struct {
    uint32_t j;
    uint32_t w[5]; /* per try_phase 4..8: kind<<29 | key(24 bit) */
} orc_record;

/* ---- src/crc.rs ---- */
const uint32_t *orc_crc_table(void);                                  /* :3-260  */
uint32_t orc_modes_checksum(const uint8_t *msg, size_t bits);          /* :263-282 */

/* ---- src/icao_filter.rs ---- */
void orc_filter_flush(orc_filter *f);                                  /* :11-17 */
uint32_t orc_icao_hash(uint32_t a);                                    /* :19-43 */
void orc_filter_add(orc_filter *f, uint32_t addr);                     /* :46-62 */
int orc_filter_test(const orc_filter *f, uint32_t addr);               /* :65-97 */

/* ---- src/mode_s/mod.rs ---- */
size_t orc_getbits(const uint8_t *data, size_t first_1idx, size_t last_1idx); /* :14-30 */
/* returns 1 for Some((msglen,score)), 0 for None */
int orc_score_modes_message(orc_filter *f, const uint8_t *msg, size_t msg_bytes,
                            int *msglen_bytes, int *score);            /* :34-139 */

/* ---- src/utils.rs ---- */
/* iq: interleaved (re, im) int16 pairs in memory order (num_complex::Complex<i16>);
 * returns 0, or -1 when n > 131072 (the reference panics, src/lib.rs:48). */
int orc_to_mag(const int16_t *iq_re_im, size_t n, orc_magbuf *out);    /* :43-58 */
uint16_t orc_mag_one(int16_t re, int16_t im);                          /* :47-55 */
/* not reference behaviour: stream continuity (checker for B200ADSB_OPT_CARRY) */
int orc_to_mag_carry(const int16_t *prev_tail326, const int16_t *iq_re_im, size_t n, orc_magbuf *out);

/* ---- src/demod_2400.rs ---- */
/* returns 1 and fills high/sig/noise on Some, else 0 */
int orc_check_preamble(const uint16_t *p14, int32_t *high, uint32_t *base_signal,
                       uint32_t *base_noise);                          /* :215-321 */
/* slices the 14 message bytes of one try_phase; data = &mag.data[0] */
void orc_slice_phase(const uint16_t *data, size_t j, int try_phase, uint8_t *msg14); /* :158-182 */
/* passes the preamble + SNR + quiet-zone gates?  (:127-146) */
int orc_gate(const uint16_t *data, size_t j);
/* full demodulator; returns number of frames (may exceed cap: extra ones dropped) */
size_t orc_demodulate2400(orc_filter *f, const orc_magbuf *mag, orc_frame *out,
                          size_t cap);                                 /* :115-212 */

/* ---- order-free two-pass form, used to check sharded runs ---- */
uint32_t orc_classify(const uint8_t *msg14);                           /* A.5 */
size_t orc_demod_records(const orc_magbuf *mag, orc_record *out, size_t cap);

/* ---- convenience: the reference bench routine (benches/demod_benchmark.rs:7-12) ---- */
size_t orc_routine(orc_filter *f, const int16_t *iq_re_im, size_t n, orc_frame *out,
                   size_t cap, int flush_first);
/* runs `iters` repetitions of the routine over n_buffers buffers on `threads`
 * threads (each thread: private filter, private buffers round-robin); returns
 * elapsed seconds and total frames through *frames. */
double orc_bench(const int16_t *iq_re_im, size_t n_buffers, size_t samples_per_buffer,
                 int iters, int threads, int flush_each, uint64_t *frames);

#ifdef __cplusplus
}
#endif
#endif
