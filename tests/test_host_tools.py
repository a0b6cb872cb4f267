"""The host-side pieces around the hot path that SURVEY.md §8(f) ranks next: the AVR/TCP emitter
(dump1090_rs/src/main.rs:149-200) and the receive loop over a file / pipe source."""
import os
import re
import socket
import subprocess
import time

import numpy as np
import pytest

from conftest import REPO, oracle_stream

HOST = os.path.join(REPO, "dump1090_rs_b200", "host")


def _build(name, extra=()):
    exe = os.path.join(HOST, name)
    src = exe + ".cpp"
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(
            os.path.getmtime(os.path.join(HOST, f)) for f in os.listdir(HOST) if f.endswith((".cpp", ".hpp"))):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, src, *extra])
    return exe


def test_avr_server_semantics():
    """AvrServer alone (no CUDA): non-blocking accept, complete lines to every client in order, a
    client that resets its connection is dropped (main.rs:154-157,183-199)."""
    exe = _build("avr_server_test")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    assert "avr_server_test ok" in out.stdout


def _start_receiver(args):
    exe = _build("dump1090_b200", ["-L" + os.path.join(REPO, "dump1090_rs_b200"), "-lb200adsb", "-Wl,-rpath,$ORIGIN/.."])
    p = subprocess.Popen([exe, "--port", "0", "--wait-clients", "1", *args], stdin=subprocess.PIPE,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    line = p.stderr.readline().decode()
    m = re.search(r"listening on [\d.]+:(\d+)", line)
    assert m, line
    s = socket.create_connection(("127.0.0.1", int(m.group(1))), timeout=30)
    return p, s


def _drain(sock, p):
    sock.settimeout(60)
    data = b""
    while True:
        chunk = sock.recv(65536)
        if not chunk:
            break
        data += chunk
    sock.close()
    p.wait(timeout=60)
    return data.decode().split()


@pytest.mark.gpu
def test_receiver_loop_on_capture_file(tmp_path, captures, golden_frames):
    """dump1090_b200 --file: the reference binary's loop on a capture in the reference's on-disk
    format; the TCP client and stdout both get the golden AVR lines of tests/test.rs."""
    from dump1090_rs_b200 import utils
    name = "test_1641427457780"
    path = utils.save_test_data(captures[name], str(tmp_path / (name + ".iq")))
    p, s = _start_receiver(["--file", path])
    p.stdin.close()
    got = _drain(s, p)
    want = ["*" + g["hex"] + ";" for g in golden_frames[name]]
    assert got == want
    assert [l for l in p.stdout.read().decode().split("\n") if l.startswith("*")] == want
    assert p.returncode == 0


@pytest.mark.gpu
@pytest.mark.parametrize("batch", [1, 3])
def test_receiver_loop_on_raw_pipe(batch, captures, oracle_mod):
    """dump1090_b200 --raw -: raw CS16 (re, im) from a pipe in reads of 50,000 samples (ragged last
    read), alone or three reads per call: the lines are the oracle's for the same sequence of reads."""
    mtu = 50000
    iq = np.concatenate([captures[k] for k in sorted(captures)])          # 3 x 131072 samples
    reads = [iq[i:i + mtu] for i in range(0, len(iq), mtu)]
    ref, _ = oracle_stream(oracle_mod, reads)
    p, s = _start_receiver(["--raw", "-", "--mtu", str(mtu), "--batch", str(batch), "--quiet"])
    p.stdin.write(np.ascontiguousarray(iq).tobytes())
    p.stdin.close()
    got = _drain(s, p)
    assert got == ["*" + f["msg"].hex() + ";" for f in ref]
    assert len(got) > 0
    assert b"*" not in p.stdout.read()                                    # --quiet: no frames on stdout
    assert p.returncode == 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["double", "serial"])
def test_receiver_loop_double_buffer_and_recovery(mode, oracle_mod):
    """The double-buffered loop (read k+1 overlaps the demodulation of read k, main.rs:161-167) emits what
    the serial loop and the oracle emit -- also when a queued batch overflows the candidate pool (a read
    that is a dense periodic pattern) and has to be redone together with the batch queued behind it."""
    from dump1090_rs_b200 import synth
    n = 131072
    reads = [synth.make_buffer(55, k, msgs_per_buffer=25, icao_pool=5)[0] for k in range(7)]
    pat = np.array([3, 8, 11, 4, 11, 1, 4, 9, 8, 1, 11])
    dense = reads[3].copy()
    dense[:, 0] = (250 * np.tile(pat, n // len(pat) + 1)[:n] + 100).astype(np.int16)
    dense[:, 1] = 0
    reads[3] = dense
    ref, _ = oracle_stream(oracle_mod, reads)
    assert len(ref) > 30
    args = ["--raw", "-", "--quiet", "--frame-cap", "4096"] + (["--serial"] if mode == "serial" else [])
    p, s = _start_receiver(args)
    p.stdin.write(np.ascontiguousarray(np.concatenate(reads)).tobytes())
    p.stdin.close()
    got = _drain(s, p)
    assert got == ["*" + f["msg"].hex() + ";" for f in ref]
    assert p.returncode == 0


def test_cli_and_config_layering(tmp_path):
    """main.rs:33-120 without a device: defaults (--driver rtlsdr, embedded gains), --driver-extra appended to
    the SoapySDR argument string, --custom-config entries pushed to the front one by one (the LAST entry of
    the file is found first), exact driver match, an unknown driver is an error.  No GPU needed."""
    exe = _build("dump1090_b200", ["-L" + os.path.join(REPO, "dump1090_rs_b200"), "-lb200adsb", "-Wl,-rpath,$ORIGIN/.."])
    run = lambda *a: subprocess.run([exe, "--print-config", *a], capture_output=True, text=True, timeout=30)
    r = run()
    assert r.returncode == 0
    assert "[-] using soapysdr driver_args: driver=rtlsdr\n" in r.stdout
    assert "[-] Writing gain: TUNER = 49.6" in r.stdout
    r = run("--driver", "hackrf", "--driver-extra", "serial=1234", "--driver-extra", "bias=1")
    assert "driver_args: driver=hackrf,serial=1234,bias=1" in r.stdout
    assert "Writing gain: LNA = 40" in r.stdout and "Writing gain: VGA = 52" in r.stdout
    r = run("--driver", "uhd")
    assert "setting antenna: RX2" in r.stdout and "Writing gain: PGA = 70" in r.stdout
    cfg = tmp_path / "custom.toml"
    cfg.write_text('''# user overrides
[[sdrs]]
driver = "rtlsdr"   # first entry: shadowed by the later one
[[sdrs.gain]]
key = "TUNER"
value = 10.0

[[sdrs]]
driver = "rtlsdr"
channel = 1
[[sdrs.setting]]
key = "biastee"
value = "true"
[[sdrs.gain]]
key = "TUNER"
value = 20.5

[[sdrs]]
driver = "airspy"
[[sdrs.gain]]
key = "LNA"
value = 3
''')
    r = run("--custom-config", str(cfg))
    assert r.returncode == 0
    assert f"[-] read in custom config: {cfg}" in r.stdout
    assert "channel 1" in r.stdout and "Writing gain: TUNER = 20.5" in r.stdout and "TUNER = 10" not in r.stdout
    assert "Writing setting: biastee = true" in r.stdout
    assert "Writing gain: LNA = 3" in run("--custom-config", str(cfg), "--driver", "airspy").stdout
    assert "Writing gain: LNA = 40" in run("--custom-config", str(cfg), "--driver", "hackrf").stdout    # embedded list still there
    r = run("--driver", "rtl")                     # exact match only
    assert r.returncode == 1 and "not found in custom or default config" in r.stderr
    bad = tmp_path / "bad.toml"
    bad.write_text('[[sdrs]]\ndriver = "x"\n')    # `gain` is required (sdrconfig.rs:15)
    assert run("--custom-config", str(bad)).returncode == 1
