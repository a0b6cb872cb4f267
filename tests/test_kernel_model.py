"""CPU checks of the algebra the CUDA scan kernel relies on (kernels.cuh), against the oracle:
closed-form bit positions, difference form of the correlators, edge-bit form of the preamble
templates, and the CRC-24 field tables.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from dump1090_rs_b200 import _ffi

W = {0: (5, -3, -2, 0), 1: (4, -1, -3, 0), 2: (3, 1, -4, 0), 3: (2, 3, -5, 0), 4: (1, 5, -5, -1)}


def test_closed_form_bit_positions(captures, oracle_mod):
    """demod_2400.rs:158-182 == bit n of try_phase t at P = 5(j+19)+t+12n (sample P//5,
    correlator P%5), and correlators on first differences (kernels.cuh P2)."""
    o = oracle_mod.Oracle()
    mb = o.to_mag(captures["test_1641427457780"])
    d = oracle_mod.mag_array(mb).astype(np.int64)
    L = oracle_mod.lib()
    rng = np.random.default_rng(1)
    msg = (C.c_uint8 * 14)()
    for j in list(rng.integers(0, 131072 - 1, 40)) + [0, 131071, 21915]:
        for t in range(4, 9):
            L.orc_slice_phase(mb.data, int(j), t, msg)
            ref = np.unpackbits(np.frombuffer(bytes(msg), dtype=np.uint8))
            n = np.arange(112)
            P = 5 * (int(j) + 19) + t + 12 * n
            i, phi = P // 5, P % 5
            got = np.zeros(112, dtype=np.uint8)
            for k in range(112):
                a, b, c, e = d[i[k]:i[k] + 4]
                u, v, w = a - b, b - c, c - e
                x = [5 * u + 2 * v, 4 * u + 3 * v, 3 * u + 4 * v, 2 * u + 5 * v, u + 6 * v + w][phi[k]]
                wt = W[int(phi[k])]
                assert x == wt[0] * a + wt[1] * b + wt[2] * c + wt[3] * e
                got[k] = x > 0
            assert (got == ref).all(), (j, t)
            assert i.max() + 3 <= int(j) + 290


def test_template_edge_bits(oracle_mod):
    """check_preamble (demod_2400.rs:215-321) == AND of rising/falling edge bits."""
    rng = np.random.default_rng(2)
    L = oracle_mod.lib()
    hits = 0
    for trial in range(20000):
        p = rng.integers(0, 40, 14).astype(np.uint16) if trial % 2 else rng.integers(0, 65536, 14).astype(np.uint16)
        R = [int(p[k] < p[k + 1]) for k in range(13)]
        F = [int(p[k] > p[k + 1]) for k in range(13)]
        quick = R[0] & F[12]
        T3 = F[1] & R[2] & F[3] & R[8] & F[9] & R[10]
        T4 = F[1] & R[2] & F[3] & R[8] & F[9] & R[11]
        T5 = F[1] & R[2] & F[4] & R[8] & F[10] & R[11]
        T6 = F[1] & R[3] & F[4] & R[9] & F[10] & R[11]
        T7 = F[2] & R[3] & F[4] & R[9] & F[10] & R[11]
        hi, sg, ns = C.c_int32(), C.c_uint32(), C.c_uint32()
        ok = L.orc_check_preamble(p.ctypes.data, C.byref(hi), C.byref(sg), C.byref(ns))
        assert bool(ok) == bool(quick & (T3 | T4 | T5 | T6 | T7))
        if ok:
            hits += 1
            q = p.astype(np.int64)
            if T3:
                exp = ((q[1] + q[3] + q[9] + q[11] + q[12]) // 4, q[1] + q[3] + q[9], q[5] + q[6] + q[7])
            elif T4:
                exp = ((q[1] + q[3] + q[9] + q[12]) // 4, q[1] + q[3] + q[9] + q[12], q[5] + q[6] + q[7] + q[8])
            elif T5:
                exp = ((q[1] + q[3] + q[4] + q[9] + q[10] + q[12]) // 4, q[1] + q[12], q[6] + q[7])
            elif T6:
                exp = ((q[1] + q[4] + q[10] + q[12]) // 4, q[1] + q[4] + q[10] + q[12], q[5] + q[6] + q[7] + q[8])
            else:
                exp = ((q[1] + q[2] + q[4] + q[10] + q[12]) // 4, q[4] + q[10] + q[12], q[6] + q[7] + q[8])
            assert (hi.value, sg.value, ns.value) == tuple(int(x) for x in exp)
    assert hits > 100


def _fields(msg14: bytes):
    bits = np.unpackbits(np.frombuffer(msg14, dtype=np.uint8))
    f = [0] * 5
    for n in range(112):
        if bits[n]:
            f[n % 5] |= 1 << (n // 5)
    return f


def _mulx(s):
    s <<= 1
    return s ^ 0x1FFF409 if s & 0x1000000 else s


def test_crc_field_tables(oracle_mod):
    """kernels.cuh syn112_fields/syn56_fields with the tables built in b200adsb.cu ==
    modes_checksum (crc.rs:263-282)."""
    L = _ffi.lib()
    tabs = np.zeros(840 + 256, dtype=np.uint32)
    assert L.b200adsb_debug_crc_tabs(tabs.ctypes.data) == 840
    assert (tabs[840:] == oracle_mod.crc_table()).all()
    t = [int(x) for x in tabs[:840]]
    a112 = lambda f: t[f & 0xFF] ^ t[256 + ((f >> 8) & 0xFF)] ^ t[512 + ((f >> 16) & 0x3F)]
    a56 = lambda f: t[576 + (f & 0xFF)] ^ t[576 + 256 + ((f >> 8) & 7)]
    rng = np.random.default_rng(3)
    msgs = [bytes(rng.integers(0, 256, 14, dtype=np.uint8)) for _ in range(3000)]
    msgs += [bytes([0] * k + [1 << b] + [0] * (13 - k)) for k in range(14) for b in range(8)]
    for m in msgs:
        f = _fields(m)
        s = a112(f[0])
        for r in range(1, 5):
            s = _mulx(s) ^ a112(f[r])
        s ^= (((f[0] >> 22) & 1) << 1) ^ ((f[1] >> 22) & 1)
        assert s == oracle_mod.modes_checksum(m, 112)
        s = a56(f[0])
        for r in range(1, 5):
            s = _mulx(s) ^ a56(f[r])
        s ^= (f[0] >> 11) & 1
        assert s == oracle_mod.modes_checksum(m[:7], 56)
        # address bits 8..31 from the fields (msg_bits<8,24>)
        v = 0
        for n in range(8, 32):
            v = (v << 1) | ((f[n % 5] >> (n // 5)) & 1)
        assert v == int.from_bytes(m[1:4], "big")


def test_abi_exports_match_header():
    """Every function declared in include/b200adsb.h is exported by libb200adsb.so."""
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(_ffi.__file__), "..", "include", "b200adsb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(b200adsb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    L = _ffi.lib()
    for n in sorted(names):
        assert hasattr(L, n), n
    assert names == set(_ffi.EXPORTS)
    assert L.b200adsb_version() >= 100
    # pure host helper: Jenkins hash of icao_filter.rs:19-43
    from oracle import oracle as O
    for a in (0, 1, 0xABCDEF, 0xFFFFFF, 0x2ABCDEF):
        assert L.b200adsb_icao_hash(a) == O.icao_hash(a)


def test_avr_lines(golden_frames):
    """main.rs:174-176: "*{hex};\\n" per frame (host formatting, no GPU)."""
    from dump1090_rs_b200 import avr
    name = "test_1641427457780"
    frames = [dict(msg=bytes.fromhex(g["hex"])) for g in golden_frames[name]]
    txt = avr.format_frames(frames)
    assert txt == "".join("*%s;\n" % g["hex"] for g in golden_frames[name]).encode()
    assert avr.format_frames([]) == b""
