"""The file-replay front-end (apps/wibeth_tpg_algorithms_emulator.cpp; SURVEY.md 8f-3, docs/README.md:20-48,74-88): a raw .bin
of concatenated 7200-byte frames goes through WIBEthFrameProcessor (the C++ shim) and comes out as the reference emulator's
trigger-primitive text file. CPU: the binary exists, documents the reference's options and refuses what it cannot do.
GPU: a generated frame file, replayed, gives exactly the oracle's TPs (and the unpacked ADC dump)."""
import os
import subprocess

import numpy as np
import pytest

import fdreadoutlibs_b200 as S
from fdreadoutlibs_b200 import frames as F
from oracle import binding as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "build", "bin", "wibeth_tpg_algorithms_emulator")


def run(*args, cwd=None):
    return subprocess.run([EMU, *args], capture_output=True, text=True, cwd=cwd, timeout=300)


def test_emulator_options_match_the_reference_emulator():
    assert os.path.exists(EMU), "build it with `make apps` (or __graft_entry__.build())"
    r = run("--help")
    assert r.returncode == 0
    for opt in ("-f,--frame-file-path", "-a,--algorithm", "-i,--implementation", "-d,--duration-test", "-n,--num-frames-to-read",
                "-t,--tpg-threshold", "--save-adc-data", "--save-trigprim"):  # docs/README.md:29-40
        assert opt in r.stdout, opt
    assert run().returncode == 2                                          # no frame file
    r = run("-f", "/nonexistent.bin")
    assert r.returncode == 1 and "cannot open" in r.stderr
    r = run("-f", EMU, "-i", "NAIVE")                                      # no CPU implementation in this library
    assert r.returncode == 2 and "no CPU implementation" in r.stderr


@pytest.mark.skipif(S.device_available(), reason="this check is for GPU-less hosts")
def test_emulator_fails_loudly_without_a_gpu(tmp_path):
    fr = S.gen_wibeth_host(S.gen_params(3, 0.2), 1, 4)
    path = tmp_path / "frames.bin"
    fr.tofile(path)
    r = run("-f", str(path), "-t", "30")
    assert r.returncode == 1 and "no CUDA device" in r.stderr  # no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("algorithm,thr,links", [("SimpleThreshold", 30, 1), ("AbsRS", 40, 1), ("SimpleThreshold", 45, 3)])
def test_replay_of_a_frame_file_matches_the_oracle(tmp_path, algorithm, thr, links):
    n = 150
    fr = S.gen_wibeth_host(S.gen_params(71, 0.3), 1, n)[0]
    # the file a readout application would replay: stream id 0 of crate 1, slot 0 (the generator's geo id)
    path = tmp_path / "frames.bin"
    fr.tofile(path)
    r = run("-f", str(path), "-a", algorithm, "-i", "AVX", "-t", str(thr), "-n", "130", "--save-trigprim", "--save-adc-data", "--links", str(links),
            "--superchunk", "16", "--out-prefix", str(tmp_path / "run"))
    assert r.returncode == 0, r.stderr
    assert "Read 130 frames" in r.stdout
    got = np.loadtxt(tmp_path / "run_trigprim.txt", delimiter=",", skiprows=1, dtype=np.uint64, ndmin=2)
    cfg = B.make_config(algorithm=S.ALGORITHMS[algorithm], threshold=thr)
    want = F.sort_tps(B.Oracle(cfg).process(fr[:130]))
    assert want.size > 50
    assert f"Found {links * want.size} hits" in r.stdout
    # link 0 replays the file as it is: offline channel of the "linear" stand-in map = ((crate*8+slot)*64+stream)*64 + chan, looked up
    # the way production does (H2: the position-ordered LUT indexed with the frame channel)
    crate, slot, stream = 1, 0, 0
    lut = np.array([((crate * 8 + slot) * 64 + stream) * 64 + ((p & ~15) | F.LANE_PERM[p & 15]) for p in range(64)])
    first = got[: want.size]
    order = np.lexsort((lut[want["channel"]], want["time_start"]))
    assert (first[:, 0] == lut[want["channel"]][order]).all()
    for col, field in ((1, "time_start"), (2, "time_over_threshold"), (3, "time_peak"), (4, "adc_integral"), (5, "adc_peak")):
        assert (first[:, col] == want[field][order].astype(np.uint64)).all(), field
    assert (first[:, 6] == 1).all()  # Type::kTPC
    if links > 1:  # replicated links: same hits, their own stream id in the offline channel
        second = got[want.size: 2 * want.size]
        assert (second[:, 1:] == first[:, 1:]).all() and (second[:, 0] == first[:, 0] + 64).all()
    adc = np.loadtxt(tmp_path / "run_adc_data.txt", delimiter=",", dtype=np.uint16)
    ref_adc, _ = F.unpack_wibeth_frames(fr[:130])
    assert (adc == ref_adc.reshape(-1, 64)).all()
