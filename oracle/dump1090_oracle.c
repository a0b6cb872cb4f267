/*
 * oracle/dump1090_oracle.c -- CPU restatement of the dump1090_rs hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see dump1090_oracle.h).  Parity: PINNED against the
 * reference's golden vectors (tests/test.rs) by tests/test_oracle_golden.py.
 *
 * This file restates, function by function, what the reference computes; it is
 * written from the behaviour of the cited lines, not transliterated.  The CRC
 * table is generated from the Mode-S generator polynomial rather than copied
 * (tests/test_oracle_golden.py checks it against src/crc.rs:3-260 when the
 * reference tree is mounted).
 */
#define _GNU_SOURCE
#include "dump1090_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ crc.rs */

static uint32_t g_crc_table[256];
static pthread_once_t g_crc_once = PTHREAD_ONCE_INIT;

/* src/crc.rs:3-260: 256-entry table of the 24-bit Mode-S CRC, generator
 * 0x1FFF409 (entry 1 is 0x00fff409, src/crc.rs:5), MSB-first. */
static void crc_table_init(void)
{
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i << 16;
        for (int k = 0; k < 8; k++)
            c = (c & 0x800000u) ? ((c << 1) ^ 0xFFF409u) : (c << 1);
        g_crc_table[i] = c & 0xFFFFFFu;
    }
}

const uint32_t *orc_crc_table(void)
{
    pthread_once(&g_crc_once, crc_table_init);
    return g_crc_table;
}

/* src/crc.rs:263-282 */
uint32_t orc_modes_checksum(const uint8_t *msg, size_t bits)
{
    const uint32_t *t = orc_crc_table();
    size_t n = bits / 8;
    uint32_t rem = 0;
    for (size_t i = 0; i + 3 < n; i++) {
        rem = (rem << 8) ^ t[msg[i] ^ ((rem & 0xff0000u) >> 16)];
        rem &= 0xffffffu;
    }
    rem ^= ((uint32_t)msg[n - 3] << 16) ^ ((uint32_t)msg[n - 2] << 8) ^ (uint32_t)msg[n - 1];
    return rem;
}

/* ---------------------------------------------------------- icao_filter.rs */

/* src/icao_filter.rs:11-17 */
void orc_filter_flush(orc_filter *f)
{
    memset(f->a, 0, sizeof f->a);
    memset(f->b, 0, sizeof f->b);
    f->full_events = 0;
}

/* src/icao_filter.rs:19-43: Jenkins one-at-a-time over the three low bytes,
 * carried in 64-bit arithmetic, truncated to 32 bits, masked to 12 bits. */
uint32_t orc_icao_hash(uint32_t a32)
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
    return (uint32_t)h & (ORC_ICAO_FILTER_SIZE - 1);
}

/* src/icao_filter.rs:46-62 */
void orc_filter_add(orc_filter *f, uint32_t addr)
{
    uint32_t h = orc_icao_hash(addr), h0 = h;
    while (f->a[h] != 0 && f->a[h] != addr) {
        h = (h + 1) & (ORC_ICAO_FILTER_SIZE - 1);
        if (h == h0) {
            f->full_events++; /* reference: eprintln!("icao24 hash table full") */
            return;
        }
    }
    if (f->a[h] == 0)
        f->a[h] = addr;
}

/* src/icao_filter.rs:65-97 */
int orc_filter_test(const orc_filter *f, uint32_t addr)
{
    uint32_t h0 = orc_icao_hash(addr), h = h0;
    while (f->a[h] != 0 && f->a[h] != addr) {
        h = (h + 1) & (ORC_ICAO_FILTER_SIZE - 1);
        if (h == h0)
            break;
    }
    if (f->a[h] == addr)
        return 1;
    h = h0;
    while (f->b[h] != 0 && f->b[h] != addr) {
        h = (h + 1) & (ORC_ICAO_FILTER_SIZE - 1);
        if (h == h0)
            break;
    }
    return f->b[h] == addr;
}

/* -------------------------------------------------------------- mode_s/mod.rs */

/* src/mode_s/mod.rs:14-30 (1-indexed, inclusive, MSB first) */
size_t orc_getbits(const uint8_t *data, size_t first_1idx, size_t last_1idx)
{
    size_t ans = 0;
    for (size_t bit = first_1idx - 1; bit <= last_1idx - 1; bit++)
        ans = (ans << 1) | ((data[bit >> 3] >> (7 - (bit & 7))) & 1u);
    return ans;
}

/* src/mode_s/mod.rs:34-139 */
int orc_score_modes_message(orc_filter *f, const uint8_t *msg, size_t msg_bytes,
                            int *msglen_bytes, int *score)
{
    size_t validbits = msg_bytes * 8;
    if (validbits < ORC_SHORT_MSG_BYTES * 8) /* :37-39 */
        return 0;
    unsigned df = (unsigned)orc_getbits(msg, 1, 5); /* :41 */
    size_t msgbits = (df & 0x10) ? ORC_LONG_MSG_BYTES * 8 : ORC_SHORT_MSG_BYTES * 8;
    if (validbits < msgbits) /* :48-50 */
        return 0;
    int all_zero = 1; /* :51-53 */
    for (size_t i = 0; i < msg_bytes; i++)
        if (msg[i]) {
            all_zero = 0;
            break;
        }
    if (all_zero)
        return 0;

    int res;
    switch (df) {
    case 0: case 4: case 5: { /* :56-72 */
        uint32_t crc = orc_modes_checksum(msg, msgbits);
        res = orc_filter_test(f, crc) ? 1000 : -1;
        break;
    }
    case 11: { /* :73-90 */
        uint32_t crc = orc_modes_checksum(msg, msgbits);
        uint32_t iid = crc & 0x7f;
        crc &= 0xffff80u;
        uint32_t addr = (uint32_t)orc_getbits(msg, 9, 32);
        int member = orc_filter_test(f, addr);
        if (crc != 0)
            res = -2;
        else if (iid == 0 && member)
            res = 1600;
        else if (iid == 0) {
            orc_filter_add(f, addr);
            res = 750;
        } else
            res = member ? 1000 : -1;
        break;
    }
    case 17: case 18: { /* :91-109 */
        uint32_t addr = (uint32_t)orc_getbits(msg, 9, 32);
        uint32_t crc = orc_modes_checksum(msg, msgbits);
        int member = orc_filter_test(f, addr);
        if (crc != 0)
            res = -2;
        else if (member)
            res = 1800;
        else {
            orc_filter_add(f, df == 17 ? addr : (addr | ORC_ICAO_FILTER_ADSB_NT));
            res = 1400;
        }
        break;
    }
    case 16: case 20: case 21: /* :110-119 */
    case 24: case 25: case 26: case 27: case 28: case 29: case 30: case 31: { /* :120-134 */
        uint32_t crc = orc_modes_checksum(msg, ORC_LONG_MSG_BYTES * 8);
        res = orc_filter_test(f, crc) ? 1000 : -2;
        break;
    }
    default: /* :135 */
        res = -2;
    }
    *msglen_bytes = (df & 0x10) ? ORC_LONG_MSG_BYTES : ORC_SHORT_MSG_BYTES;
    *score = res;
    return 1;
}

/* -------------------------------------------------------------------- utils.rs */

/* src/utils.rs:47-55.  `i` is the imaginary part and the FMA multiplicand,
 * the real part is squared and rounded first; f32 throughout; Rust `as u16`
 * saturates. */
uint16_t orc_mag_one(int16_t re, int16_t im)
{
    float fi = (float)im / 32768.0f;
    float fq = (float)re / 32768.0f;
    volatile float q2 = fq * fq; /* force the separate rounding (no contraction) */
    float mag_sqr = fmaf(fi, fi, q2);
    float mag = sqrtf(mag_sqr);
    float v = fmaf(mag, 65535.0f, 0.5f);
    if (v >= 65535.0f)
        return 65535;
    return (uint16_t)v; /* v >= 0 always */
}

/* src/utils.rs:43-58 + src/lib.rs:36-51 */
int orc_to_mag(const int16_t *iq, size_t n, orc_magbuf *out)
{
    if (n > ORC_MAG_BUF_SAMPLES)
        return -1;
    memset(out->data, 0, sizeof out->data); /* MagnitudeBuffer::default(), lib.rs:36-44 */
    for (size_t k = 0; k < n; k++)
        out->data[ORC_TRAILING_SAMPLES + k] = orc_mag_one(iq[2 * k], iq[2 * k + 1]);
    out->length = n;
    return 0;
}

/* Stream-continuity variant (NOT reference behaviour, which zero-fills: utils.rs:44,
 * lib.rs:36-50): the 326 leading slots hold the magnitudes of the previous buffer's last
 * samples, as the C dump1090 the crate was translated from did.  prev_tail: 326 (re, im)
 * pairs, oldest first, or NULL.  Checker for B200ADSB_OPT_CARRY. */
int orc_to_mag_carry(const int16_t *prev_tail, const int16_t *iq, size_t n, orc_magbuf *out)
{
    if (orc_to_mag(iq, n, out) != 0)
        return -1;
    if (prev_tail)
        for (size_t k = 0; k < ORC_TRAILING_SAMPLES; k++)
            out->data[k] = orc_mag_one(prev_tail[2 * k], prev_tail[2 * k + 1]);
    return 0;
}

/* ---------------------------------------------------------------- demod_2400.rs */

/* src/demod_2400.rs:215-321 */
int orc_check_preamble(const uint16_t *p, int32_t *high, uint32_t *sig, uint32_t *noise)
{
    if (!(p[0] < p[1] && p[12] > p[13])) /* :221 */
        return 0;
    if (p[1] > p[2] && p[2] < p[3] && p[3] > p[4] && p[8] < p[9] && p[9] > p[10] &&
        p[10] < p[11]) { /* :226-241, peaks 1,3,9,11-12 */
        *high = ((int32_t)p[1] + p[3] + p[9] + p[11] + p[12]) / 4;
        *sig = (uint32_t)p[1] + p[3] + p[9];
        *noise = (uint32_t)p[5] + p[6] + p[7];
        return 1;
    }
    if (p[1] > p[2] && p[2] < p[3] && p[3] > p[4] && p[8] < p[9] && p[9] > p[10] &&
        p[11] < p[12]) { /* :242-262, peaks 1,3,9,12 */
        *high = ((int32_t)p[1] + p[3] + p[9] + p[12]) / 4;
        *sig = (uint32_t)p[1] + p[3] + p[9] + p[12];
        *noise = (uint32_t)p[5] + p[6] + p[7] + p[8];
        return 1;
    }
    if (p[1] > p[2] && p[2] < p[3] && p[4] > p[5] && p[8] < p[9] && p[10] > p[11] &&
        p[11] < p[12]) { /* :263-279, peaks 1,3-4,9-10,12 */
        *high = ((int32_t)p[1] + p[3] + p[4] + p[9] + p[10] + p[12]) / 4;
        *sig = (uint32_t)p[1] + p[12];
        *noise = (uint32_t)p[6] + p[7];
        return 1;
    }
    if (p[1] > p[2] && p[3] < p[4] && p[4] > p[5] && p[9] < p[10] && p[10] > p[11] &&
        p[11] < p[12]) { /* :280-300, peaks 1,4,10,12 */
        *high = ((int32_t)p[1] + p[4] + p[10] + p[12]) / 4;
        *sig = (uint32_t)p[1] + p[4] + p[10] + p[12];
        *noise = (uint32_t)p[5] + p[6] + p[7] + p[8];
        return 1;
    }
    if (p[2] > p[3] && p[3] < p[4] && p[4] > p[5] && p[9] < p[10] && p[10] > p[11] &&
        p[11] < p[12]) { /* :301-317, peaks 1-2,4,10,12 */
        *high = ((int32_t)p[1] + p[2] + p[4] + p[10] + p[12]) / 4;
        *sig = (uint32_t)p[4] + p[10] + p[12];
        *noise = (uint32_t)p[6] + p[7] + p[8];
        return 1;
    }
    return 0;
}

/* src/demod_2400.rs:127-146 */
int orc_gate(const uint16_t *data, size_t j)
{
    int32_t high;
    uint32_t sig, noise;
    if (!orc_check_preamble(&data[j], &high, &sig, &noise))
        return 0;
    if (sig * 2 < 3 * noise) /* :129 */
        return 0;
    static const int quiet[9] = {5, 6, 7, 8, 14, 15, 16, 17, 18}; /* :135-143 */
    for (int k = 0; k < 9; k++)
        if ((int32_t)data[j + quiet[k]] >= high)
            return 0;
    return 1;
}

/* The five correlators of src/demod_2400.rs:72-83, the index advance of
 * :62-68 and the phase walk of :50-58 / byte restart of :38-46, stated as
 * small tables indexed by the phase number 0..4. */
static inline int32_t phase_bit(int ph, const uint16_t *m)
{
    switch (ph) {
    case 0: return 5 * (int32_t)m[0] - 3 * (int32_t)m[1] - 2 * (int32_t)m[2];
    case 1: return 4 * (int32_t)m[0] - (int32_t)m[1] - 3 * (int32_t)m[2];
    case 2: return 3 * (int32_t)m[0] + (int32_t)m[1] - 4 * (int32_t)m[2];
    case 3: return 2 * (int32_t)m[0] + 3 * (int32_t)m[1] - 5 * (int32_t)m[2];
    default: return (int32_t)m[0] + 5 * (int32_t)m[1] - 5 * (int32_t)m[2] - (int32_t)m[3];
    }
}
static const int k_advance[5] = {2, 2, 2, 3, 3};     /* :62-68 */
static const int k_next[5] = {2, 3, 4, 0, 1};        /* :50-58: 0->2->4->1->3->0 */

/* src/demod_2400.rs:158-182 */
void orc_slice_phase(const uint16_t *data, size_t j, int try_phase, uint8_t *msg)
{
    size_t slice_loc = j + 19 + (size_t)(try_phase / 5); /* :159 */
    int phase = try_phase % 5;                           /* :160 */
    for (int b = 0; b < ORC_LONG_MSG_BYTES; b++) {
        const uint16_t *s = &data[slice_loc];
        int start = phase;
        unsigned byte = 0;
        size_t index = 0;
        for (int i = 0; i < 8; i++) {
            if (phase_bit(phase, &s[index]) > 0)
                byte |= 1u << (7 - i);
            index += (size_t)k_advance[phase];
            phase = k_next[phase];
        }
        msg[b] = (uint8_t)byte;
        slice_loc += index;
        phase = (start + 1) % 5; /* next_start, :38-46 */
    }
}

/* src/demod_2400.rs:115-212 */
size_t orc_demodulate2400(orc_filter *f, const orc_magbuf *mag, orc_frame *out, size_t cap)
{
    const uint16_t *data = mag->data;
    size_t n_out = 0;
    for (size_t j = 0; j < mag->length; j++) {
        if (!orc_gate(data, j))
            continue;
        orc_frame best;
        memset(&best, 0, sizeof best);
        best.score = -2; /* :152 */
        best.len = ORC_SHORT_MSG_BYTES;
        uint8_t msg[ORC_LONG_MSG_BYTES];
        for (int t = 4; t < 9; t++) { /* :158 */
            orc_slice_phase(data, j, t, msg);
            int len, score;
            if (orc_score_modes_message(f, msg, ORC_LONG_MSG_BYTES, &len, &score) &&
                score > best.score) { /* :184-185 */
                best.len = (uint8_t)len;
                memcpy(best.msg, msg, ORC_LONG_MSG_BYTES);
                best.score = score;
                best.phase = (uint8_t)t;
                uint64_t p = 0; /* :191-198 (unobservable through buffer()) */
                size_t signal_len = ORC_LONG_MSG_BYTES * 12 / 5;
                for (size_t k = 0; k < signal_len; k++) {
                    uint64_t m = data[j + 19 + k];
                    p += m * m;
                }
                best.signal_level = (double)p / 65535.0 / 65535.0 / (double)signal_len;
            }
        }
        if (best.score < 0) /* :203 */
            continue;
        best.j = (uint32_t)j;
        if (n_out < cap)
            out[n_out] = best;
        n_out++;
    }
    return n_out;
}

/* ------------------------------------------------------- two-pass (order-free) */

/* what score_modes_message would do, as a function of filter
 * membership only.  Returns kind<<29 | key. */
uint32_t orc_classify(const uint8_t *msg)
{
    int all_zero = 1;
    for (int i = 0; i < ORC_LONG_MSG_BYTES; i++)
        if (msg[i])
            all_zero = 0;
    if (all_zero)
        return 0;
    unsigned df = msg[0] >> 3;
    uint32_t addr = ((uint32_t)msg[1] << 16) | ((uint32_t)msg[2] << 8) | msg[3];
    uint32_t kind = ORC_K_NONE, key = 0;
    switch (df) {
    case 0: case 4: case 5:
        kind = ORC_K_PAR_SHORT;
        key = orc_modes_checksum(msg, 56);
        break;
    case 11: {
        uint32_t crc = orc_modes_checksum(msg, 56);
        if ((crc & 0xffff80u) == 0) {
            kind = (crc & 0x7f) ? ORC_K_DF11_IID : ORC_K_DF11_IID0;
            key = addr;
        }
        break;
    }
    case 17: case 18:
        if (orc_modes_checksum(msg, 112) == 0) {
            kind = df == 17 ? ORC_K_DF17 : ORC_K_DF18;
            key = addr;
        }
        break;
    case 16: case 20: case 21:
    case 24: case 25: case 26: case 27: case 28: case 29: case 30: case 31:
        kind = ORC_K_PAR_LONG;
        key = orc_modes_checksum(msg, 112);
        break;
    default:
        break;
    }
    return (kind << 29) | key;
}

size_t orc_demod_records(const orc_magbuf *mag, orc_record *out, size_t cap)
{
    size_t n = 0;
    uint8_t msg[ORC_LONG_MSG_BYTES];
    for (size_t j = 0; j < mag->length; j++) {
        if (!orc_gate(mag->data, j))
            continue;
        orc_record r;
        r.j = (uint32_t)j;
        for (int t = 4; t < 9; t++) {
            orc_slice_phase(mag->data, j, t, msg);
            r.w[t - 4] = orc_classify(msg);
        }
        if (n < cap)
            out[n] = r;
        n++;
    }
    return n;
}

/* --------------------------------------------------------------- bench helpers */

/* benches/demod_benchmark.rs:7-12 */
size_t orc_routine(orc_filter *f, const int16_t *iq, size_t n, orc_frame *out, size_t cap,
                   int flush_first)
{
    static __thread orc_magbuf *mb = NULL;
    if (!mb)
        mb = (orc_magbuf *)malloc(sizeof *mb);
    if (flush_first)
        orc_filter_flush(f);
    if (orc_to_mag(iq, n, mb) != 0)
        return 0;
    return orc_demodulate2400(f, mb, out, cap);
}

typedef struct {
    const int16_t *iq;
    size_t n_buffers, spb;
    int iters, tid, threads, flush_each;
    uint64_t frames;
} bench_arg;

static void *bench_thread(void *p)
{
    bench_arg *a = (bench_arg *)p;
    orc_filter *f = (orc_filter *)malloc(sizeof *f);
    orc_magbuf *mb = (orc_magbuf *)malloc(sizeof *mb);
    orc_frame *fr = (orc_frame *)malloc(4096 * sizeof *fr);
    orc_filter_flush(f);
    uint64_t frames = 0;
    for (int it = 0; it < a->iters; it++)
        for (size_t b = (size_t)a->tid; b < a->n_buffers; b += (size_t)a->threads) {
            if (a->flush_each)
                orc_filter_flush(f);
            orc_to_mag(a->iq + 2 * b * a->spb, a->spb, mb);
            frames += orc_demodulate2400(f, mb, fr, 4096);
        }
    a->frames = frames;
    free(fr);
    free(mb);
    free(f);
    return NULL;
}

double orc_bench(const int16_t *iq, size_t n_buffers, size_t spb, int iters, int threads,
                 int flush_each, uint64_t *frames)
{
    if (threads < 1)
        threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    bench_arg *args = (bench_arg *)malloc(sizeof(bench_arg) * (size_t)threads);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < threads; i++) {
        args[i] = (bench_arg){iq, n_buffers, spb, iters, i, threads, flush_each, 0};
        pthread_create(&th[i], NULL, bench_thread, &args[i]);
    }
    uint64_t total = 0;
    for (int i = 0; i < threads; i++) {
        pthread_join(th[i], NULL);
        total += args[i].frames;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (frames)
        *frames = total;
    free(th);
    free(args);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
