"""Parity of the CUDA path (through the C ABI) against the CPU oracle.  Bit-exact: the
path is integer/byte work plus one f32 magnitude whose rounding is specified
(src/utils.rs:47-55), so the tolerance is zero everywhere."""
import numpy as np
import pytest

from conftest import frames_key, oracle_stream

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import dump1090_rs_b200 as d
    return d


@pytest.fixture()
def ctx(pkg):
    c = pkg.Context(0)
    yield c
    c.close()


NAMES = ["test_1641427457780", "test_1641428165033", "test_1641428106243"]


# ---------------------------------------------------------------- reference's own tests
@pytest.mark.parametrize("name", NAMES)
def test_reference_routine_on_captures(name, pkg, ctx, captures, golden_frames, oracle_mod):
    """tests/test.rs:7-17 through the mirrored API: icao_flush, to_mag, demodulate2400."""
    pkg.icao_filter.icao_flush(ctx)
    outbuf = pkg.utils.to_mag(captures[name], ctx)
    data = pkg.demod_2400.demodulate2400(outbuf, ctx)
    gold = [bytes.fromhex(g["hex"]) for g in golden_frames[name]]
    for a, b in zip(data, gold):
        assert a.buffer() == b
    # full list against the oracle (count, j, phase, score, bytes)
    o = oracle_mod.Oracle()
    ref = o.demod_iq(captures[name], flush=True)
    assert [(m.j, m.phase, m.score, m.buffer().hex()) for m in data] == \
        [(f["j"], f["phase"], f["score"], f["msg"].hex()) for f in ref]
    assert len(data) in (5, 6)


@pytest.mark.parametrize("name", NAMES)
def test_to_mag_bit_exact_on_captures(name, pkg, ctx, captures, oracle_mod):
    outbuf = pkg.utils.to_mag(captures[name], ctx)
    ref = oracle_mod.mag_array(oracle_mod.Oracle().to_mag(captures[name]))
    assert outbuf.length == 131072
    assert np.array_equal(outbuf.data, ref)


def test_fused_equals_two_step(pkg, ctx, captures):
    for name in NAMES:
        ctx.icao_flush()
        a = frames_key(ctx.demod_iq(captures[name]))
        ctx.icao_flush()
        d, n = ctx.to_mag(captures[name])
        b = frames_key(ctx.demodulate2400(d, n))
        assert a == b and len(a) >= 5


def test_stream_of_captures_persistent_filter(ctx, captures, oracle_mod):
    """Three buffers as one stream (no flush in between) == sequential reference, both as
    three calls and as one batched call."""
    bufs = [captures[n] for n in NAMES]
    ref, o = oracle_stream(oracle_mod, bufs)
    got = []
    for b, iq in enumerate(bufs):
        for f in ctx.demod_iq(iq):
            f["buffer"] = b
            got.append(f)
    assert frames_key(got) == frames_key(ref)
    assert set(ctx.icao_snapshot()) == o.members()
    ctx.icao_flush()
    batch = np.stack(bufs)
    got2, counts = ctx.demod_iq_batch(batch, 3, 131072, want_counts=True)
    assert frames_key(got2) == frames_key(ref)
    assert list(counts) == [sum(1 for f in ref if f["buffer"] == b) for b in range(3)]


@pytest.mark.parametrize("tile", [8, 32, 88, 352, 472, 1024, 4096, 7768, 8184])
def test_tile_sizes(tile, pkg, captures, oracle_mod):
    from dump1090_rs_b200 import _ffi
    c = pkg.Context(0)
    c.set_option(_ffi.OPT_TILE, tile)
    ref = oracle_mod.Oracle().demod_iq(captures[NAMES[2]], flush=True)
    assert frames_key(c.demod_iq(captures[NAMES[2]])) == frames_key(ref)
    c.close()


# ---------------------------------------------------------------- to_mag arithmetic
def test_to_mag_random_full_range(ctx, oracle_mod):
    rng = np.random.default_rng(7)
    iq = rng.integers(-32768, 32768, (131072, 2), dtype=np.int64).astype(np.int16)
    iq[:8] = [[0, 0], [32767, 32767], [-32768, -32768], [-32768, 32767], [1, 0], [0, 1], [-1, -1], [23170, 23170]]
    d, n = ctx.to_mag(iq)
    ref = oracle_mod.mag_array(oracle_mod.Oracle().to_mag(iq))
    assert np.array_equal(d, ref)
    # small amplitudes (sqrt near the denormal-free low end)
    iq2 = rng.integers(-40, 41, (65536, 2)).astype(np.int16)
    d2, n2 = ctx.to_mag(iq2)
    assert n2 == 65536
    assert np.array_equal(d2, oracle_mod.mag_array(oracle_mod.Oracle().to_mag(iq2)))


def test_fast_magnitude_exhaustive(ctx):
    """The scan kernel's magnitude (no I2F/F2I, rsqrt-seeded sqrt) equals the IEEE statement of
    utils.rs:47-55 on every one of the 2^32 (re, im) inputs; the latter is checked against the
    oracle by test_to_mag_random_full_range (to_mag_kernel uses it)."""
    import ctypes as C
    from dump1090_rs_b200 import _ffi
    m, f = C.c_uint64(123), C.c_uint32(0)
    rc = _ffi.lib().b200adsb_debug_mag_sweep(ctx._h, C.byref(m), C.byref(f))
    assert rc == 0
    assert m.value == 0, f"{m.value} mismatches, first at re|im<<16 = {f.value:#010x}"


# ---------------------------------------------------------------- synthetic streams
@pytest.mark.parametrize("msgs", [0, 1, 10, 100])
def test_synthetic_stream(msgs, ctx, oracle_mod):
    """BASELINE configs 3/4 at oracle-sized batches: rtl-like noise + injected DF17."""
    from dump1090_rs_b200 import synth
    nb = 12
    batch = synth.make_batch(1090, nb, msgs_per_buffer=msgs)
    ref, o = oracle_stream(oracle_mod, list(batch))
    got = ctx.demod_iq_batch(batch, nb, 131072)
    assert frames_key(got) == frames_key(ref)
    assert set(ctx.icao_snapshot()) == o.members()
    if msgs >= 10:
        assert len(got) > msgs * nb // 4


def test_full_range_noise(ctx, oracle_mod):
    from dump1090_rs_b200 import synth
    bufs = [synth.full_range_buffer(5, b) for b in range(4)]
    ref, _ = oracle_stream(oracle_mod, bufs)
    got = ctx.demod_iq_batch(np.stack(bufs), 4, 131072)
    assert frames_key(got) == frames_key(ref)


# ---------------------------------------------------------------- edge cases
def test_empty_and_tiny_inputs(ctx, oracle_mod, captures):
    assert ctx.demod_iq(np.zeros((0, 2), dtype=np.int16)) == []
    d, n = ctx.to_mag(np.zeros((0, 2), dtype=np.int16))
    assert n == 0 and not d.any()
    for n in (1, 13, 326, 327, 1000):
        iq = captures[NAMES[0]][21000:21000 + n]
        ref = oracle_mod.Oracle().demod_iq(iq)
        assert frames_key(ctx.demod_iq(iq)) == frames_key(ref)
    with pytest.raises(IndexError):
        ctx.to_mag(np.zeros((131073, 2), dtype=np.int16))


def test_ragged_batch_and_stride(ctx, captures, oracle_mod):
    """Buffers of different lengths in one call (lengths[]), with a stride larger than the
    payload; a frame near the end of a short buffer must vanish exactly like in the
    reference (the last 326 samples are never scanned, lib.rs:24,47-50)."""
    base = captures[NAMES[0]]
    lens = [131072, 22300, 21915 - 326 + 200, 70000, 5]
    stride = 131080
    batch = np.zeros((len(lens), stride, 2), dtype=np.int16)
    bufs = []
    for b, ln in enumerate(lens):
        batch[b, :ln] = base[:ln]
        batch[b, ln:] = 12345        # garbage beyond the payload must be ignored
        bufs.append(base[:ln])
    ref, _ = oracle_stream(oracle_mod, bufs)
    got = ctx.demod_iq_batch(batch, len(lens), 131072, stride=stride, lengths=lens)
    assert frames_key(got) == frames_key(ref)


def test_pool_growth(pkg, captures, oracle_mod):
    """A candidate pool that is too small is grown and the batch redone, exactly."""
    from dump1090_rs_b200 import _ffi
    c = pkg.Context(0)
    c.set_option(_ffi.OPT_POOL_SHIFT, 16)
    bufs = [captures[n] for n in NAMES] * 3
    ref, _ = oracle_stream(oracle_mod, bufs)
    got = c.demod_iq_batch(np.stack(bufs), len(bufs), 131072)
    assert frames_key(got) == frames_key(ref)
    c.close()


def test_output_capacity_error(pkg, ctx, captures):
    from dump1090_rs_b200 import _ffi
    with pytest.raises(pkg.B200AdsbError) as e:
        ctx.demod_iq(captures[NAMES[0]], cap=2)
    assert e.value.status == _ffi.ERR_CAPACITY


def test_filter_preload_and_snapshot(pkg, ctx, captures, oracle_mod):
    """A filter loaded before the stream changes scores (1400 -> 1800) like the reference."""
    o = oracle_mod.Oracle()
    for a in (0xAD9293, 0xAA2BC4, 0x123456):
        o.icao_filter_add(a)
        ctx.icao_filter_add(a)
    assert ctx.icao_filter_test(0xAD9293) and not ctx.icao_filter_test(0x654321) and ctx.icao_filter_test(0)
    ref = o.demod_iq(captures[NAMES[0]])
    got = ctx.demod_iq(captures[NAMES[0]])
    assert frames_key(got) == frames_key(ref)
    assert any(f["score"] == 1800 for f in got)
    snap = ctx.icao_snapshot()
    assert set(snap) == o.members()
    c2 = pkg.Context(0)
    c2.icao_restore(snap)
    assert set(c2.icao_snapshot()) == set(snap)
    assert frames_key(c2.demod_iq(captures[NAMES[1]])) == frames_key(o.demod_iq(captures[NAMES[1]]))
    c2.close()


# ---------------------------------------------------------------- crc.rs / mode_s
def test_modes_checksum(ctx, oracle_mod, golden_frames):
    rng = np.random.default_rng(11)
    msgs = rng.integers(0, 256, (4096, 14), dtype=np.uint8)
    for bits in (56, 112):
        got = ctx.modes_checksum(msgs, bits)
        ref = [oracle_mod.modes_checksum(bytes(m[: bits // 8]), bits) for m in msgs]
        assert list(got) == ref


def test_score_modes_messages_sequence(ctx, oracle_mod):
    """score_modes_message over a sequence with adds and later membership hits, every DF."""
    import ctypes as C
    from dump1090_rs_b200 import synth
    rng = np.random.default_rng(13)
    msgs = []
    for i in range(3000):
        m = bytearray(rng.integers(0, 256, 14, dtype=np.uint8))
        m[0] = (int(rng.integers(0, 32)) << 3) | (m[0] & 7)
        msgs.append(bytes(m))
    good = [synth.df17_message(0xA00000 + 17 * k, bytes(rng.integers(0, 256, 7, dtype=np.uint8))) for k in range(40)]
    for k, g in enumerate(good):          # valid DF17s, repeated so later ones score 1800
        msgs.insert(50 * k + 7, g)
        msgs.insert(50 * k + 31, g)
    df18 = bytearray(good[3]); df18[0] = (18 << 3) | 5
    body = bytes(df18[:11]); p = synth._crc24(body); df18[11:] = bytes([(p >> 16) & 255, (p >> 8) & 255, p & 255])
    msgs += [bytes(df18), bytes(df18), bytes(14)]
    # DF11 with syndrome 0 for a fresh address, twice (750 then 1600), and DF4 hitting a member
    df11 = bytearray([11 << 3, 0xA0, 0x00, 0x99, 0, 0, 0]); p = synth._crc24(bytes(df11[:4]))
    df11[4:7] = bytes([(p >> 16) & 255, (p >> 8) & 255, p & 255])
    msgs += [bytes(df11) + bytes(7), bytes(df11) + bytes(7)]
    arr = np.frombuffer(b"".join(msgs), dtype=np.uint8).reshape(-1, 14)
    lens, scores = ctx.score_modes_messages(arr)
    o = oracle_mod.Oracle()
    L = oracle_mod.lib()
    for i, m in enumerate(msgs):
        ln, sc = C.c_int(0), C.c_int(0)
        buf = (C.c_uint8 * 14).from_buffer_copy(m)
        ok = L.orc_score_modes_message(C.byref(o.filter), buf, 14, C.byref(ln), C.byref(sc))
        if not ok:
            assert lens[i] == 0, i
        else:
            assert (int(lens[i]), int(scores[i])) == (ln.value, sc.value), (i, m.hex())
    assert set(ctx.icao_snapshot()) == o.members()
    assert (scores == 1800).sum() >= 40 and (scores == 1600).sum() >= 1


def test_filter_capacity_rule(pkg, oracle_mod):
    """More than 4096 distinct addresses: only the first 4096 by first-add order enter the
    filter ("icao24 hash table full", icao_filter.rs:50-55)."""
    import ctypes as C
    from dump1090_rs_b200 import synth
    c = pkg.Context(0)
    msgs = [synth.df17_message(0x400000 + k, bytes([k & 255] * 7)) for k in range(4200)]
    msgs = msgs + msgs[4000:4200] + msgs[:50]
    arr = np.frombuffer(b"".join(msgs), dtype=np.uint8).reshape(-1, 14)
    lens, scores = c.score_modes_messages(arr)
    o = oracle_mod.Oracle()
    L = oracle_mod.lib()
    ref = []
    for m in msgs:
        ln, sc = C.c_int(0), C.c_int(0)
        buf = (C.c_uint8 * 14).from_buffer_copy(m)
        L.orc_score_modes_message(C.byref(o.filter), buf, 14, C.byref(ln), C.byref(sc))
        ref.append(sc.value)
    assert list(scores) == ref
    assert len(c.icao_snapshot()) == 4096 and set(c.icao_snapshot()) == o.members()
    c.close()


# ---------------------------------------------------------------- device-resident batch
def test_device_batch_and_properties(pkg, oracle_mod):
    """BASELINE config-3 sized batch resident on the device (torch only as the allocator):
    size-independent properties + oracle parity on a sample of buffers."""
    import torch
    from dump1090_rs_b200 import synth, _ffi
    nb = 256
    iq = synth.noise_batch_torch(1090, nb)
    frames = torch.zeros((4096, 28), dtype=torch.uint8, device="cuda")
    counts = torch.zeros(nb, dtype=torch.int32, device="cuda")
    c = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
    n = c.demod_iq_batch_ptr(iq.data_ptr(), nb, 131072, 131072, frames.data_ptr(), 4096,
                             counts_ptr=counts.data_ptr())
    torch.cuda.synchronize()
    assert int(counts.sum()) == n
    raw = frames[:n].cpu().numpy()
    key = raw[:, 24:28].copy().view(np.uint32)[:, 0].astype(np.int64) * (1 << 20) + \
        raw[:, 20:24].copy().view(np.uint32)[:, 0]
    assert (np.diff(key) >= 0).all()                     # ordered by (buffer, j), duplicates kept
    # idempotence: same batch on a fresh context gives the same bytes
    c2 = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
    frames2 = torch.zeros_like(frames)
    n2 = c2.demod_iq_batch_ptr(iq.data_ptr(), nb, 131072, 131072, frames2.data_ptr(), 4096)
    assert n2 == n and torch.equal(frames[:n], frames2[:n])
    # oracle on the first buffers of the stream (the filter state is a prefix property)
    host = iq[:6].cpu().numpy()
    ref, _ = oracle_stream(oracle_mod, list(host))
    got = [dict(buffer=int(r[24:28].view(np.uint32)[0]), j=int(r[20:24].view(np.uint32)[0]), phase=int(r[15]),
                score=int(r[16:18].view(np.int16)[0]), msg=bytes(r[: r[14]])) for r in raw if r[24:28].view(np.uint32)[0] < 6]
    assert frames_key(got) == frames_key(ref)
    c.close(); c2.close()


def test_split_scan_resolve_two_shards(pkg, oracle_mod):
    """The multi-GPU protocol on one GPU: two contexts take alternate buffers, exchange their
    ICAO add-events, and together reproduce the single-stream result exactly."""
    import torch
    from dump1090_rs_b200 import synth
    nb = 8
    batch = synth.make_batch(77, nb, msgs_per_buffer=40, icao_pool=8)
    ref, o = oracle_stream(oracle_mod, list(batch))
    shards = [np.ascontiguousarray(batch[r::2]) for r in range(2)]
    ctxs = [pkg.Context(0) for _ in range(2)]
    dev = [torch.from_numpy(s).cuda() for s in shards]
    pairs = [torch.zeros((4096, 2), dtype=torch.int64, device="cuda") for _ in range(2)]
    cnt = []
    for r in range(2):
        ctxs[r].scan_batch_dev(dev[r].data_ptr(), nb // 2, 131072, 131072, r, 2)
        cnt.append(ctxs[r].events_export_dev(pairs[r].data_ptr(), 4096))
    assert sum(cnt) > 0
    got = []
    for r in range(2):
        ctxs[r].events_import_dev(pairs[1 - r].data_ptr(), cnt[1 - r])
        out = torch.zeros((4096, 28), dtype=torch.uint8, device="cuda")
        n = ctxs[r].resolve_batch_dev(out.data_ptr(), 4096)
        for rr in out[:n].cpu().numpy():
            got.append(dict(buffer=int(rr[24:28].view(np.uint32)[0]) * 2 + r, j=int(rr[20:24].view(np.uint32)[0]),
                            phase=int(rr[15]), score=int(rr[16:18].view(np.int16)[0]), msg=bytes(rr[: rr[14]])))
    got.sort(key=lambda f: (f["buffer"], f["j"]))
    assert frames_key(got) == frames_key(ref)
    for r in range(2):
        assert set(ctxs[r].icao_snapshot()) == o.members()
        ctxs[r].close()


def test_cpp_host_mirror_reference_routine(tmp_path, captures, golden_frames):
    """The C++ mirror of the crate surface (dump1090_rs_b200/host/dump1090_rs.hpp) running the
    reference's own routine (tests/test.rs:7-17) on a capture file in the reference's on-disk
    format, compared with the golden vectors."""
    import os
    import subprocess
    from dump1090_rs_b200 import utils
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(here, "dump1090_rs_b200", "host")
    exe = os.path.join(host, "reference_routine")
    if not os.path.exists(exe):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(host, "reference_routine.cpp"),
                               "-L" + os.path.join(here, "dump1090_rs_b200"), "-lb200adsb", "-Wl,-rpath,$ORIGIN/.."])
    name = NAMES[2]
    path = utils.save_test_data(captures[name], str(tmp_path / (name + ".iq")))
    assert np.array_equal(utils.read_test_data(path), captures[name])
    out = subprocess.run([exe, path], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["*" + g["hex"] + ";" for g in golden_frames[name]]


def test_packed_event_exchange_two_shards(pkg, oracle_mod):
    """The sync-free form of the exchange (row 0 = count) used by sharded.ShardedDemodulator:
    two contexts on one GPU, the 'all-gather' done by hand."""
    import torch
    from dump1090_rs_b200 import synth
    nb, rows = 6, 512
    batch = synth.make_batch(91, nb, msgs_per_buffer=25, icao_pool=5)
    ref, o = oracle_stream(oracle_mod, list(batch))
    ctxs = [pkg.Context(0) for _ in range(2)]
    dev = [torch.from_numpy(np.ascontiguousarray(batch[r::2])).cuda() for r in range(2)]
    gathered = torch.zeros((2 * rows, 2), dtype=torch.int64, device="cuda")
    for r in range(2):
        ctxs[r].scan_batch_dev(dev[r].data_ptr(), nb // 2, 131072, 131072, r, 2)
        ctxs[r].events_pack_dev(gathered[r * rows:].data_ptr(), rows)
    torch.cuda.synchronize()
    assert int(gathered[0, 0]) + int(gathered[rows, 0]) > 0
    got = []
    for r in range(2):
        ctxs[r].events_import_packed_dev(gathered.data_ptr(), 2, rows, r)
        out = torch.zeros((4096, 28), dtype=torch.uint8, device="cuda")
        n = ctxs[r].resolve_batch_dev(out.data_ptr(), 4096)
        for rr in out[:n].cpu().numpy():
            got.append(dict(buffer=int(rr[24:28].view(np.uint32)[0]) * 2 + r, j=int(rr[20:24].view(np.uint32)[0]),
                            phase=int(rr[15]), score=int(rr[16:18].view(np.int16)[0]), msg=bytes(rr[: rr[14]])))
    got.sort(key=lambda f: (f["buffer"], f["j"]))
    assert frames_key(got) == frames_key(ref)
    for c in ctxs:
        assert set(c.icao_snapshot()) == o.members()
        c.close()


def test_randomized_ragged_streams(pkg, captures, oracle_mod):
    """Many small random streams: random buffer counts, lengths (0..9000), strides, tile sizes
    and contents (capture slices with real frames, injected DF17s, full-range noise); every one
    must equal the sequential reference, including the filter at the end."""
    from dump1090_rs_b200 import _ffi, synth
    rng = np.random.default_rng(2024)
    base = np.concatenate([captures[n] for n in NAMES])
    dense, _ = synth.make_buffer(5, 0, msgs_per_buffer=120, icao_pool=4)
    for case in range(24):
        nb = int(rng.integers(1, 6))
        spb = int(rng.integers(1, 9000))
        stride = spb + int(rng.integers(0, 3)) * 4 + (0 if rng.random() < 0.5 else int(rng.integers(0, 7)))
        lens = [int(rng.integers(0, spb + 1)) for _ in range(nb)]
        batch = np.full((nb, stride, 2), 777, dtype=np.int16)
        bufs = []
        for b in range(nb):
            kind = rng.integers(0, 3)
            if kind == 0:      # slice of a real capture around a known frame
                j0 = int(rng.choice([21915, 68286, 71134, 130601, 131072 + 14611, 262144 + 9323])) - 326 - int(rng.integers(0, 300))
                src = base[max(j0, 0): max(j0, 0) + lens[b]]
            elif kind == 1:    # dense synthetic traffic
                o0 = int(rng.integers(0, 131072 - spb))
                src = dense[o0:o0 + lens[b]]
            else:
                src = synth.full_range_buffer(case, b, n=max(lens[b], 1))[:lens[b]]
            lens[b] = min(lens[b], len(src))
            batch[b, :lens[b]] = src[:lens[b]]
            bufs.append(batch[b, :lens[b]].copy())
        ref, o = oracle_stream(oracle_mod, bufs)
        c = pkg.Context(0)
        if rng.random() < 0.6:
            c.set_option(_ffi.OPT_TILE, int(rng.choice([8, 24, 88, 472, 856, 2008, 7384])))
        got = c.demod_iq_batch(batch, nb, spb, stride=stride, lengths=lens)
        assert frames_key(got) == frames_key(ref), (case, nb, spb, stride, lens)
        assert set(c.icao_snapshot()) == o.members(), case
        c.close()


def test_carry_option_stream_continuity(pkg, oracle_mod):
    """B200ADSB_OPT_CARRY: bit-exact against the oracle's carry variant, buffer by buffer and in
    one batched call, and invariant under where the stream is cut."""
    from dump1090_rs_b200 import _ffi, synth
    stream, _ = synth.make_buffer(31, 0, n=100000, msgs_per_buffer=90, icao_pool=5)
    # rotate the stream so that a decodable message straddles the cut at sample 25000
    whole = oracle_mod.Oracle().demod_iq_carry(stream)
    stream = np.ascontiguousarray(np.roll(stream, -((whole[12]["j"] - 326 + 100) - 25000), axis=0))
    cuts = [0, 25000, 50000, 75000, 100000]
    o = oracle_mod.Oracle()
    ref = []
    for b in range(4):
        for f in o.demod_iq_carry(stream[cuts[b]:cuts[b + 1]]):
            f["buffer"] = b
            ref.append(f)
    c = pkg.Context(0)
    c.set_option(_ffi.OPT_CARRY, 1)
    got = []
    for b in range(4):                                   # four calls
        for f in c.demod_iq(stream[cuts[b]:cuts[b + 1]]):
            f["buffer"] = b
            got.append(f)
    assert frames_key(got) == frames_key(ref)
    c.icao_flush()
    c.set_option(_ffi.OPT_CARRY, 1)                      # restart continuity
    batch = np.ascontiguousarray(stream.reshape(4, 25000, 2))
    assert frames_key(c.demod_iq_batch(batch, 4, 25000)) == frames_key(ref)   # one batched call
    # cut somewhere else: same frames at the same stream positions
    c.icao_flush()
    c.set_option(_ffi.OPT_CARRY, 1)
    pos = lambda fr, starts: [(starts[f["buffer"]] + f["j"], f["msg"].hex()) for f in fr]
    cuts2 = [0, 10123, 10500, 60001, 100000]
    got2 = []
    for b in range(4):
        for f in c.demod_iq(stream[cuts2[b]:cuts2[b + 1]]):
            f["buffer"] = b
            got2.append(f)
    assert pos(got2, cuts2) == pos(ref, cuts)
    # and more than the reference semantics find
    c.set_option(_ffi.OPT_CARRY, 0)
    c.icao_flush()
    assert len(c.demod_iq_batch(batch, 4, 25000)) < len(ref)
    c.close()


@pytest.mark.gpu
def test_enqueue_only_batches_match_synchronous_calls(pkg):
    """b200adsb_demod_iq_batch_dev_async: several batches queued back to back on one context give
    the frames, counts and filter state of the same batches through the synchronous call."""
    import torch
    from dump1090_rs_b200 import synth
    nbuf, spb, nbatch, cap = 24, 131072, 4, 4096
    iq = synth.make_batch(77, nbuf * nbatch, msgs_per_buffer=6).reshape(nbatch, nbuf, spb, 2)
    dev = torch.device("cuda", 0)
    d_iq = torch.from_numpy(np.ascontiguousarray(iq)).to(dev)
    stream = torch.cuda.current_stream()

    def run(queued):
        ctx = pkg.Context(0, stream.cuda_stream)
        ctx.icao_flush()
        frames = torch.zeros((nbatch, cap, 28), dtype=torch.uint8, device=dev)
        res = torch.zeros((nbatch, 4), dtype=torch.int32, device=dev)
        counts = []
        if not queued:
            for k in range(nbatch):
                counts.append(ctx.demod_iq_batch_ptr(d_iq[k].data_ptr(), nbuf, spb, spb, frames[k].data_ptr(), cap))
        else:
            # size the candidate pool once, as a caller's warm-up would; then queue everything
            ctx.demod_iq_batch_ptr(d_iq[0].data_ptr(), nbuf, spb, spb, frames[0].data_ptr(), cap)
            ctx.icao_flush()
            for k in range(nbatch):
                ctx.demod_iq_batch_async_ptr(d_iq[k].data_ptr(), nbuf, spb, spb, frames[k].data_ptr(), cap,
                                             res[k].data_ptr())
            ctx.sync()
            r = res.cpu().numpy()
            assert (r[:, 1] == 0).all() and (r[:, 3] == 0).all(), r
            counts = [int(x) for x in r[:, 0]]
        snap = sorted(ctx.icao_snapshot())
        out = [frames[k, :counts[k]].cpu().numpy().tobytes() for k in range(nbatch)]
        ctx.close()
        return counts, out, snap

    c_sync, f_sync, s_sync = run(False)
    c_q, f_q, s_q = run(True)
    assert sum(c_sync) > 0
    assert c_q == c_sync
    assert f_q == f_sync
    assert s_q == s_sync


def test_dense_template_matches_overflow_the_match_list(pkg, ctx, oracle_mod):
    """A periodic magnitude pattern on which 3 of every 11 positions match a preamble template AND
    pass the SNR / quiet-zone gates (~2000 matches and survivors per tile: twice the capacity of the
    kernel's match list, 16 decode windows per tile, a candidate pool that has to grow): frames and
    the number of surviving positions must be the oracle's.  The rest of the buffer is noise with
    injected messages."""
    from dump1090_rs_b200 import synth
    pat = np.array([3, 8, 11, 4, 11, 1, 4, 9, 8, 1, 11])
    n = 131072
    iq = synth.make_batch(5, 1, msgs_per_buffer=20)[0].copy()
    part = n // 4
    levels = np.tile(pat, part // len(pat) + 1)[:part]
    iq[:part, 0] = (250 * levels + 100).astype(np.int16)
    iq[:part, 1] = 0
    orc = oracle_mod.Oracle()
    mb = orc.to_mag(iq)
    n_surv = len(orc.records(mb, cap=1 << 17))
    assert n_surv > 0.2 * part            # the premise: the pattern is dense in survivors
    ref = orc.demod_iq(iq, flush=True)
    ctx.icao_flush()
    ctx.timing(reset=True)
    got = ctx.demod_iq(iq)
    assert frames_key(got) == frames_key(ref)
    assert len(ref) > 0
    assert ctx.timing()["candidates"] == n_surv


# ---------------------------------------------------------------- round 2: per-stage parity, recovery paths
def _norm_words(ws):
    """kind NONE carries no key worth comparing (the kernel marks `None` with key 1, the oracle with 0)."""
    return [0 if (w >> 29) == 0 else w for w in ws]


def _stage1_records(pkg, iq_batch, tile=None):
    """(buffer, j, words) of every position that passed the gates, through the split scan call."""
    import torch
    from dump1090_rs_b200 import _ffi
    nb, spb = iq_batch.shape[0], iq_batch.shape[1]
    c = pkg.Context(0)
    if tile:
        c.set_option(_ffi.OPT_TILE, tile)
    d_iq = torch.from_numpy(np.ascontiguousarray(iq_batch)).to("cuda:0")
    c.scan_batch_dev(d_iq.data_ptr(), nb, spb, spb, 0, 1)
    recs = c.debug_records()
    out = torch.zeros((4096, 28), dtype=torch.uint8, device="cuda:0")
    c.resolve_batch_dev(out.data_ptr(), 4096)
    c.close()
    return [(b, j, _norm_words(w)) for b, j, w in recs]


@pytest.mark.parametrize("tile", [None, 1624, 472])
def test_stage1_records_equal_oracle_on_captures(tile, pkg, captures, oracle_mod):
    """the survivor set of the preamble gates (demod_2400.rs:127-146) and
    the stateless classification of every (j, try_phase) (demod_2400.rs:158-189, mode_s/mod.rs:34-139)
    equal orc_demod_records -- including the ~1400 candidates per capture that never become frames."""
    batch = np.stack([captures[n] for n in NAMES])
    got = _stage1_records(pkg, batch, tile)
    o = oracle_mod.Oracle()
    ref = []
    for b, n in enumerate(NAMES):
        ref += [(b, j, _norm_words(w)) for j, w in o.records(o.to_mag(captures[n]))]
    assert len(ref) > 3000
    assert got == ref


def test_stage1_records_dense_pattern_and_full_range(pkg, oracle_mod):
    from dump1090_rs_b200 import synth
    pat = np.array([3, 8, 11, 4, 11, 1, 4, 9, 8, 1, 11])
    n = 131072
    a = synth.make_batch(5, 1, msgs_per_buffer=20)[0].copy()
    part = n // 8
    levels = np.tile(pat, part // len(pat) + 1)[:part]
    a[:part, 0] = (250 * levels + 100).astype(np.int16)
    a[:part, 1] = 0
    b = synth.full_range_buffer(3, 0)
    c = synth.make_batch(9, 1, msgs_per_buffer=100)[0]
    batch = np.stack([a, b, c])
    got = _stage1_records(pkg, batch)
    o = oracle_mod.Oracle()
    ref = []
    for k in range(3):
        ref += [(k, j, _norm_words(w)) for j, w in o.records(o.to_mag(batch[k]), cap=1 << 17)]
    assert len(ref) > 5000
    assert got == ref


def test_carry_with_h2d_chunks_and_short_buffers(pkg, oracle_mod):
    """Carry mode through the host batch call when the batch is scanned in several H2D chunks
    (the chunk's first buffer continues the batch, not the previous batch) and with buffers
    shorter than the 326-sample reach-back (walked through several buffers)."""
    from dump1090_rs_b200 import _ffi, synth
    stream, _ = synth.make_buffer(41, 0, n=120000, msgs_per_buffer=110, icao_pool=6)
    rng = np.random.default_rng(7)
    for case, (spb, chunk) in enumerate([(6000, 1), (6000, 3), (2500, 2)]):
        nb = 120000 // spb
        lens = [spb] * nb
        for k in rng.choice(nb, size=nb // 3, replace=False):      # ragged, some very short buffers
            lens[int(k)] = int(rng.choice([0, 1, 50, 200, 325, 326, 327, 1000, spb - 1]))
        batch = np.zeros((nb, spb, 2), dtype=np.int16)
        bufs, pos = [], 0
        for b in range(nb):
            bufs.append(stream[pos:pos + lens[b]])
            batch[b, :lens[b]] = bufs[-1]
            pos += lens[b]
        o = oracle_mod.Oracle()
        ref = []
        for b in range(nb):
            for f in o.demod_iq_carry(bufs[b]):
                f["buffer"] = b
                ref.append(f)
        assert len(ref) > 20
        c = pkg.Context(0)
        c.set_option(_ffi.OPT_CARRY, 1)
        c.set_option(_ffi.OPT_H2D_CHUNK, chunk)
        half = nb // 2                                              # two calls: the tail crosses calls too
        got = c.demod_iq_batch(batch[:half], half, spb, lengths=lens[:half])
        for f in c.demod_iq_batch(batch[half:], nb - half, spb, lengths=lens[half:]):
            f["buffer"] += half
            got.append(f)
        assert frames_key(got) == frames_key(ref), (case, spb, chunk)
        c.close()


def test_enqueue_only_overflow_is_sticky_and_recoverable(pkg, oracle_mod):
    """A queued batch that overflows the candidate pool is not committed, and neither are the batches
    queued behind it (d_result[1] bit 3): the filter stays what it was, and re-running them with the
    synchronous call reproduces the reference."""
    import torch
    from dump1090_rs_b200 import _ffi, synth
    spb, cap = 131072, 4096
    quiet = synth.make_batch(21, 2, msgs_per_buffer=8, icao_pool=4)
    later = synth.make_batch(22, 2, msgs_per_buffer=8, icao_pool=4)          # same aircraft pool
    pat = np.array([3, 8, 11, 4, 11, 1, 4, 9, 8, 1, 11])
    dense = synth.make_batch(23, 2, msgs_per_buffer=8, icao_pool=4).copy()
    levels = np.tile(pat, spb // len(pat) + 1)[:spb]
    dense[0, :, 0] = (250 * levels + 100).astype(np.int16)
    dense[0, :, 1] = 0
    batches = [quiet, dense, later]
    ref, o = oracle_stream(oracle_mod, [b for batch in batches for b in batch])
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(np.ascontiguousarray(b)).to(dev) for b in batches]
    stream = torch.cuda.current_stream()
    ctx = pkg.Context(0, stream.cuda_stream)
    frames = torch.zeros((3, cap, 28), dtype=torch.uint8, device=dev)
    res = torch.zeros((3, 4), dtype=torch.int32, device=dev)
    n0 = ctx.demod_iq_batch_ptr(d[0].data_ptr(), 2, spb, spb, frames[0].data_ptr(), cap)   # sizes the pool for quiet traffic
    snap0 = sorted(ctx.icao_snapshot())
    for k in (1, 2):
        ctx.demod_iq_batch_async_ptr(d[k].data_ptr(), 2, spb, spb, frames[k].data_ptr(), cap, res[k].data_ptr())
    ctx.sync()
    r = res.cpu().numpy()
    assert r[1, 1] & 1, r                  # the dense batch overflowed the pool
    assert r[2, 1] & 8 and r[2, 0] == 0, r   # the batch behind it was skipped
    assert sorted(ctx.icao_snapshot()) == snap0
    counts = [n0]
    for k in (1, 2):                       # recovery: the synchronous call, in order
        counts.append(ctx.demod_iq_batch_ptr(d[k].data_ptr(), 2, spb, spb, frames[k].data_ptr(), cap))
    # a later enqueue-only batch commits again
    ctx.demod_iq_batch_async_ptr(d[0].data_ptr(), 2, spb, spb, frames[0].data_ptr(), cap, res[0].data_ptr())
    ctx.sync()
    assert res.cpu().numpy()[0, 1] == 0
    got = []
    for k in range(1, 3):
        raw = frames[k, :counts[k]].cpu().numpy()
        for row in raw:
            got.append((2 * k + int(row[24:28].view(np.uint32)[0]), int(row[20:24].view(np.uint32)[0]), int(row[15]),
                        int(row[16:18].view(np.int16)[0]), bytes(row[: row[14]]).hex()))
    assert got == [t for t in frames_key(ref) if t[0] >= 2]
    assert set(ctx.icao_snapshot()) == o.members()
    ctx.close()


def test_single_message_surface(pkg, ctx, oracle_mod, golden_frames):
    """getbits / modes_checksum / score_modes_message with the reference's own shapes
    (mode_s/mod.rs:14,34, crc.rs:263)."""
    import ctypes as C
    L = oracle_mod.lib()
    rng = np.random.default_rng(3)
    msgs = [bytes.fromhex(g["hex"]) for n in NAMES for g in golden_frames[n]]
    for m in msgs:
        assert pkg.crc.modes_checksum(m, 8 * len(m), ctx) == oracle_mod.modes_checksum(m, 8 * len(m))
        for a, b in [(1, 5), (9, 32), (33, 56), (1, 1), (8 * len(m), 8 * len(m))]:
            buf = (C.c_uint8 * len(m)).from_buffer_copy(m)
            assert pkg.mode_s.getbits(m, a, b) == L.orc_getbits(buf, a, b)
    o = oracle_mod.Oracle()
    ctx.icao_flush()
    seq = msgs + [bytes(14), bytes(7), bytes([0x28]) + bytes(5), msgs[0][:7], bytes(rng.integers(0, 256, 14, dtype=np.uint8))] + msgs
    for m in seq:
        buf = (C.c_uint8 * max(len(m), 1)).from_buffer_copy(m.ljust(1, b"\0"))
        ln, sc = C.c_int(0), C.c_int(0)
        ok = L.orc_score_modes_message(C.byref(o.filter), buf, len(m), C.byref(ln), C.byref(sc))
        got = pkg.mode_s.score_modes_message(m, ctx)
        if not ok:
            assert got is None, m.hex()
        else:
            assert got is not None and (14 if got[0] == pkg.demod_2400.MsgLen.Long else 7, got[1]) == (ln.value, sc.value), m.hex()


def test_stage1_records_deep_in_a_large_batch(pkg, oracle_mod):
    """Buffers far into a large device batch (tile indices in the thousands, pool offsets in the hundred
    thousands): survivor set and record words of sampled buffers equal the oracle's -- parity of a big launch
    beyond its first buffers."""
    import torch
    from dump1090_rs_b200 import synth
    nb, spb = 384, 131072
    d_iq = synth.noise_batch_torch(7, nb, device="cuda:0")
    sample = [0, 1, 95, 200, 383]
    inj = synth.make_batch(11, len(sample), msgs_per_buffer=40)
    for k, b in enumerate(sample):
        d_iq[b] = torch.from_numpy(inj[k]).to("cuda:0")
    c = pkg.Context(0)
    c.scan_batch_dev(d_iq.data_ptr(), nb, spb, spb, 0, 1)
    bufs, rec = c.debug_records_np(cap=1 << 20)
    out = torch.zeros((1 << 15, 28), dtype=torch.uint8, device="cuda:0")
    c.resolve_batch_dev(out.data_ptr(), 1 << 15)
    c.close()
    assert len(bufs) > 300 * 1000                       # ~1400 survivors per buffer
    assert (np.diff(bufs.astype(np.int64)) >= 0).all()  # (buffer, j) order
    o = oracle_mod.Oracle()
    host = d_iq[sample].cpu().numpy()
    for k, b in enumerate(sample):
        sel = bufs == b
        got = [(int(r[0]), _norm_words([int(x) for x in r[1:]])) for r in rec[sel]]
        ref = [(j, _norm_words(w)) for j, w in o.records(o.to_mag(host[k]), cap=1 << 17)]
        assert got == ref, b


@pytest.mark.parametrize("mode", ["lengths", "carry", "unaligned"])
def test_default_tile_non_standard_batches(mode, pkg, oracle_mod):
    """The kernel form with the default tile as a compile-time constant but WITHOUT the standard-batch
    assumptions (>= 33 full-size buffers pick the default tile; per-buffer lengths, carry mode or an
    unaligned stride rule the standard form out): frames equal the oracle's."""
    from dump1090_rs_b200 import _ffi, synth
    nb, spb = 36, 131072
    batch = synth.make_batch(91, nb, msgs_per_buffer=12, icao_pool=6)
    c = pkg.Context(0)
    if mode == "lengths":
        rng = np.random.default_rng(5)
        lens = [int(x) for x in rng.integers(90000, spb + 1, nb)]
        lens[3], lens[17] = spb, 7384 * 17 + 5           # a full one, and one whose last tile is 5 positions
        ref, o = oracle_stream(oracle_mod, [batch[b, :lens[b]] for b in range(nb)])
        got = c.demod_iq_batch(batch, nb, spb, lengths=lens)
    elif mode == "carry":
        c.set_option(_ffi.OPT_CARRY, 1)
        o = oracle_mod.Oracle()
        ref = []
        for b in range(nb):
            for f in o.demod_iq_carry(batch[b]):
                f["buffer"] = b
                ref.append(f)
        got = c.demod_iq_batch(batch, nb, spb)
    else:
        stride = spb + 2                                   # stride % 4 != 0: scalar loads
        wide = np.zeros((nb, stride, 2), dtype=np.int16)
        wide[:, :spb] = batch
        ref, o = oracle_stream(oracle_mod, [batch[b] for b in range(nb)])
        import torch
        d = torch.from_numpy(wide).to("cuda:0")
        out = torch.zeros((8192, 28), dtype=torch.uint8, device="cuda:0")
        n = c.demod_iq_batch_ptr(d.data_ptr(), nb, spb, stride, out.data_ptr(), 8192)
        raw = out[:n].cpu().numpy()
        got = [dict(buffer=int(r[24:28].view(np.uint32)[0]), j=int(r[20:24].view(np.uint32)[0]), phase=int(r[15]),
                    score=int(r[16:18].view(np.int16)[0]), msg=bytes(r[: r[14]])) for r in raw]
    assert len(ref) > 100
    assert frames_key(got) == frames_key(ref), mode
    if mode != "carry":
        assert set(c.icao_snapshot()) == o.members()
    c.close()


def test_cu8_ingest_equals_cs16_path(pkg, oracle_mod):
    """Opt-in 8-bit ingest (b200adsb_demod_cu8_batch): unsigned 8-bit (I, Q) pairs expanded to CS16 on the
    device give the frames of the CS16 path -- and of the oracle -- on the host-converted samples; ragged
    stride and more buffers than one H2D chunk."""
    import ctypes as C
    from dump1090_rs_b200 import _ffi, synth
    L = _ffi.lib()
    lut = np.array([L.b200adsb_cu8_to_cs16(v) for v in range(256)], dtype=np.int16)
    nb, spb = 5, 131072
    cs16 = synth.make_batch(33, nb, msgs_per_buffer=30, icao_pool=5)
    inv = {int(v): k for k, v in enumerate(lut)}
    u8 = np.vectorize(inv.__getitem__, otypes=[np.uint8])(cs16)      # synth samples are table values
    assert (lut[u8] == cs16).all()
    ref, o = oracle_stream(oracle_mod, [cs16[b] for b in range(nb)])
    assert len(ref) > 50
    c = pkg.Context(0)
    c.set_option(_ffi.OPT_H2D_CHUNK, 2)
    assert frames_key(c.demod_cu8_batch(u8, nb, spb)) == frames_key(ref)
    assert set(c.icao_snapshot()) == o.members()
    c.icao_flush()
    assert frames_key(c.demod_iq_batch(cs16, nb, spb)) == frames_key(ref)
    # ragged: 50,001-sample buffers (device stride rounded up to 4 samples)
    c.icao_flush()
    spb2 = 50001
    cut16 = np.ascontiguousarray(cs16[:, :spb2])
    ref2, _ = oracle_stream(oracle_mod, [cut16[b] for b in range(nb)])
    assert frames_key(c.demod_cu8_batch(np.ascontiguousarray(u8[:, :spb2]), nb, spb2)) == frames_key(ref2)
    c.close()
