/*
 * oracle/orc_bench_main.c -- stand-alone timing driver of the CPU restatement
 * (test infrastructure; used only by bench.py's cpu_baseline / --impl reference legs).
 * A separate process so that the host threads are not affected by whatever the Python
 * process has loaded.
 *
 *   orc_bench <iq_file> <n_buffers> <samples_per_buffer> <iters> <threads> <flush_each>
 *
 * iq_file: raw int16 (re, im) pairs, n_buffers * samples_per_buffer samples.
 * Prints one JSON line: {"sec": s, "frames": n, "threads": t}.
 * The routine timed is the reference bench's (benches/demod_benchmark.rs:7-12):
 * [icao_flush] + to_mag + demodulate2400 per buffer.
 */
#include <stdio.h>
#include <stdlib.h>

#include "dump1090_oracle.h"

int main(int argc, char **argv)
{
    if (argc < 7) {
        fprintf(stderr, "usage: %s iq_file n_buffers spb iters threads flush_each\n", argv[0]);
        return 2;
    }
    size_t nb = (size_t)atol(argv[2]), spb = (size_t)atol(argv[3]);
    int iters = atoi(argv[4]), threads = atoi(argv[5]), flush_each = atoi(argv[6]);
    size_t n16 = nb * spb * 2;
    int16_t *iq = (int16_t *)malloc(n16 * sizeof(int16_t));
    FILE *f = fopen(argv[1], "rb");
    if (!iq || !f || fread(iq, sizeof(int16_t), n16, f) != n16) {
        fprintf(stderr, "cannot read %zu int16 from %s\n", n16, argv[1]);
        return 1;
    }
    fclose(f);
    uint64_t frames = 0;
    orc_bench(iq, nb, spb, 1, threads, flush_each, &frames); /* warm-up */
    double sec = orc_bench(iq, nb, spb, iters, threads, flush_each, &frames);
    printf("{\"sec\": %.6f, \"frames\": %llu, \"threads\": %d}\n", sec, (unsigned long long)frames, threads);
    free(iq);
    return 0;
}
