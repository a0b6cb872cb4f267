"""CPU checks of the algebra the CUDA scan kernel relies on (kernels.cuh), against the oracle:
closed-form bit positions, difference form of the correlators, edge-bit form of the preamble
templates, and the CRC-24 field tables.  No GPU needed."""
import ctypes as C

import os

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


def test_crc_lane_tables(oracle_mod):
    """kernels.cuh syn112_fields_sh/syn56_fields_sh (tables held one entry per lane, looked up by warp
    shuffle with the lane index taken modulo 32) == modes_checksum (crc.rs:263-282)."""
    L = _ffi.lib()
    lt = np.zeros(7 * 32, dtype=np.uint32)
    assert L.b200adsb_debug_crc_lane_tabs(lt.ctypes.data) == 224
    t = [[int(x) for x in lt[32 * c:32 * c + 32]] for c in range(7)]
    sh = lambda c, idx: t[c][idx & 31]          # __shfl_sync(L.t[c], idx)
    a112 = lambda f: sh(0, f) ^ sh(1, f >> 5) ^ sh(2, f >> 10) ^ sh(3, f >> 15) ^ sh(4, (f >> 20) & 3)
    a56 = lambda f: sh(5, f) ^ sh(6, f >> 5) ^ (((f >> 10) & 1) << 1)
    rng = np.random.default_rng(4)
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


def test_cu8_conversion_table():
    """b200adsb_cu8_to_cs16 (the device's cu8_expand_kernel uses the same f32 expression): SoapySDR's RTL-SDR
    u8 -> CS16 conversion, which is also the mapping that reproduces the value set of the reference's captures
    (..., -358, -102, 153, 409, ...)."""
    L = _ffi.lib()
    lut = np.array([L.b200adsb_cu8_to_cs16(v) for v in range(256)], dtype=np.int16)
    f32 = (((np.arange(256, dtype=np.float32) - np.float32(127.4)) * np.float32(1.0 / 128.0)) * np.float32(32767.0)).astype(np.int16)
    assert (lut == f32).all()
    assert {-358, -102, 153, 409} <= set(int(x) for x in lut)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "captures.npz"))
    for k in z.files:                         # every sample of the captures is a table value
        assert np.isin(np.unique(z[k]), lut).all(), k


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


# ---------------------------------------------------------------- scan7.cuh (stage-1 kernel v7)
def _case_of(p):
    """template case 0..4 of demod_2400.rs:226-317 (first match wins) or -1"""
    R = [int(p[k] < p[k + 1]) for k in range(13)]
    F = [int(p[k] > p[k + 1]) for k in range(13)]
    if not (R[0] & F[12]):
        return -1
    T = [F[1] & R[2] & F[3] & R[8] & F[9] & R[10], F[1] & R[2] & F[3] & R[8] & F[9] & R[11],
         F[1] & R[2] & F[4] & R[8] & F[10] & R[11], F[1] & R[3] & F[4] & R[9] & F[10] & R[11],
         F[2] & R[3] & F[4] & R[9] & F[10] & R[11]]
    for cs in range(5):
        if T[cs]:
            return cs
    return -1


def test_branch_free_gate_matches_oracle(oracle_mod):
    """gate_eval_bf (scan7.cuh): the five template cases as selects on the case number give the
    oracle's SNR / quiet-zone decision (demod_2400.rs:129,135-146 with :226-317)."""
    rng = np.random.default_rng(11)
    L = oracle_mod.lib()
    checked = passed = 0
    while checked < 3000:
        # a preamble-like burst so that templates match often, plus random tails
        base = rng.integers(0, 400, 32)
        for k in (1, 3, 9, 12) if rng.random() < 0.5 else (1, 4, 10, 12):
            base[k] += rng.integers(300, 3000)
        p = base.astype(np.uint16)
        cs = _case_of(p)
        if cs < 0:
            continue
        a, h, b, e, n5, n6, n7, n8 = (int(p[k]) for k in (1, 2, 3, 4, 5, 6, 7, 8))
        c, f, g, d = (int(p[k]) for k in (9, 10, 11, 12))
        bc, ef = b + c, e + f
        H = a + d + (bc if cs < 3 else 0) + (ef if cs >= 2 else 0) + (g if cs == 0 else 0) + (h if cs == 4 else 0)
        S = (a if cs < 4 else 0) + (bc if cs < 2 else 0) + (d if cs >= 1 else 0) + (ef if cs >= 3 else 0)
        N = n6 + n7 + (n5 if (0x0B >> cs) & 1 else 0) + (n8 if (0x1A >> cs) & 1 else 0)
        mx = max(n5, n6, n7, n8, *(int(p[k]) for k in (14, 15, 16, 17, 18)))
        model = (2 * S >= 3 * N) and (mx < (H >> 2))
        data = np.zeros(400, dtype=np.uint16)
        data[:32] = p
        assert bool(L.orc_gate(data.ctypes.data, 0)) == model, (cs, p[:19])
        checked += 1
        passed += model
    assert passed > 100          # both outcomes are exercised


def test_mod12_plane_templates(oracle_mod):
    """P3a of scan7.cuh: with the edge bits de-interleaved modulo 12 (plane[rho] bit q <-> index
    12q + rho), the edge at offset s of position (rho, q) is bit q + (rho+s)//12 of row
    (rho+s) % 12 -- evaluated word-parallel this gives check_preamble's decision and case."""
    rng = np.random.default_rng(12)
    n = 12 * 32 * 3
    m = rng.integers(0, 30, n + 64).astype(np.int64)
    for k in rng.integers(0, n - 20, 200):          # sprinkle preamble-like bursts
        m[k + 1] += 500; m[k + 3] += 500; m[k + 9] += 500; m[k + 12] += 400
    Rb = (m[:-1] < m[1:]).astype(np.uint64)
    Fb = (m[:-1] > m[1:]).astype(np.uint64)
    nq = (n + 64) // 12
    def plane(bits):
        rows = []
        for rho in range(12):
            v = 0
            for q in range(nq - 1):
                v |= int(bits[12 * q + rho]) << q
            rows.append(v)
        return rows
    PR, PF = plane(Rb), plane(Fb)
    mask = (1 << 64) - 1
    found = 0
    for rho in range(12):
        def X(rows, s):
            t = rho + s
            return (rows[t % 12] >> (t // 12)) & mask
        quick = X(PR, 0) & X(PF, 12)
        T3 = X(PF, 1) & X(PR, 2) & X(PF, 3) & X(PR, 8) & X(PF, 9) & X(PR, 10)
        T4 = X(PF, 1) & X(PR, 2) & X(PF, 3) & X(PR, 8) & X(PF, 9) & X(PR, 11)
        T5 = X(PF, 1) & X(PR, 2) & X(PF, 4) & X(PR, 8) & X(PF, 10) & X(PR, 11)
        T6 = X(PF, 1) & X(PR, 3) & X(PF, 4) & X(PR, 9) & X(PF, 10) & X(PR, 11)
        T7 = X(PF, 2) & X(PR, 3) & X(PF, 4) & X(PR, 9) & X(PF, 10) & X(PR, 11)
        anym = quick & (T3 | T4 | T5 | T6 | T7)
        c1, c2, c3, c4 = T4 & ~T3, T5 & ~(T3 | T4), T6 & ~(T3 | T4 | T5), ~(T3 | T4 | T5 | T6)
        b0, b1 = c1 | c3, c2 | c3
        for q in range(60):
            i = 12 * q + rho
            want = _case_of(m[i:i + 14])
            got = -1
            if (anym >> q) & 1:
                got = ((b0 >> q) & 1) | (((b1 >> q) & 1) << 1) | (((c4 >> q) & 1) << 2)
            assert got == want, (rho, q)
            found += want >= 0
    assert found > 20


def test_dense_phase_lane_mapping_and_reciprocal():
    """scan7.cuh dense phase: (group G, lane c, slot k, element e) covers every tile magnitude index
    exactly once with residue 4c+e and plane bit 16G+k; the host reciprocal used for i / Wrow is
    exact for every mask-word index."""
    NG = 40
    seen = np.zeros(NG * 192, dtype=np.int32)
    for G in range(NG):
        for c in range(3):
            for k in range(16):
                for e in range(4):
                    i = 192 * G + 12 * k + 4 * c + e
                    seen[i] += 1
                    assert i % 12 == 4 * c + e and i // 12 == 16 * G + k
    assert (seen == 1).all()
    for W in range(1, 24):
        inv = (65536 + W - 1) // W
        for i in range(12 * W):
            assert (i * inv) >> 16 == i // W


def test_host_only_entry_points(oracle_mod, golden_frames):
    """The pure-host parts of the ABI need no GPU: getbits (mode_s/mod.rs:14-30) against the oracle's on random
    messages and every (first, last) pair, icao_hash (icao_filter.rs:19-43), and the AVR line format the
    reference writes to its TCP clients (main.rs:174-176)."""
    import ctypes as C
    L, O = _ffi.lib(), oracle_mod.lib()
    rng = np.random.default_rng(8)
    for _ in range(40):
        m = bytes(rng.integers(0, 256, 14, dtype=np.uint8))
        buf = (C.c_uint8 * 14).from_buffer_copy(m)
        for a in range(1, 113, 7):
            for b in range(a, min(a + 32, 113)):
                assert L.b200adsb_getbits(m, a, b) == O.orc_getbits(buf, a, b), (m.hex(), a, b)
    for a in [0, 1, 0xABCDEF, 0x4840D6, 0xFFFFFF, (1 << 25) | 0x123456] + [int(x) for x in rng.integers(0, 1 << 24, 200)]:
        assert L.b200adsb_icao_hash(a) == O.orc_icao_hash(a)
    frames = (_ffi.Frame * 2)()
    msgs = [bytes.fromhex(golden_frames["test_1641427457780"][0]["hex"]), bytes.fromhex("02e1971ce17c84")]
    for k, m in enumerate(msgs):
        for i, byte in enumerate(m):
            frames[k].msg[i] = byte
        frames[k].len = len(m)
    out = C.create_string_buffer(128)
    n = C.c_size_t(0)
    assert L.b200adsb_format_avr(frames, 2, out, 128, C.byref(n)) == 0
    assert out.raw[: n.value].decode() == "".join("*" + m.hex() + ";\n" for m in msgs)
    assert L.b200adsb_format_avr(frames, 2, out, 10, C.byref(n)) == _ffi.ERR_CAPACITY and n.value == 2 * 3 + 2 * (14 + 7)


def test_stage1_kernel_codegen():
    """The built library carries the three compiled forms of the stage-1 kernel, and the standard-batch form
    is the code the profiles describe: packed f32x2 arithmetic, the cp.async ring, shuffle-table CRC, no
    local-memory traffic in the dense loop, and a footprint that stays inside the instruction
    cache budget the kernel was tuned for (profiles/README.md: sensitive to code size)."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        import pytest
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", _ffi.SO_PATH], capture_output=True, text=True, timeout=300).stdout
    funcs = {}
    name = None
    for line in out.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            funcs[name] = []
        elif name and line.strip().startswith("/*") and ";" in line:
            funcs[name].append(line)
    forms = {k: v for k, v in funcs.items() if "scan7_kernel" in k}
    assert len(forms) == 4, sorted(forms)          # <true,0,false> <false,0,false> <false,7384,false> <false,7384,true>
    std = next(v for k, v in forms.items() if "ILb0ELi7384ELb1" in k)
    text = "\n".join(std)
    assert 2800 < len(std) < 3200, len(std)
    for mnemonic in ("FFMA2", "FADD2", "FMUL2", "LDGSTS", "SHFL.IDX", "MUFU.RSQ", "I2F.S16", "ATOMS", "BAR.SYNC"):
        assert mnemonic in text, mnemonic
    loop = next(i for i, ln in enumerate(std) if "DEPBAR.LE" in ln)          # the dense loop starts at the ring wait
    assert not any("STL" in ln or "LDL" in ln for ln in std[loop:loop + 125]), "spills in the dense loop"
    assert "HMMA" not in text and "UTCHMMA" not in text              # no tensor-core detour on this path
