"""The whole library on a CPU, as a test: tests/simt builds wolken_b200.cu itself — C ABI, host orchestration, every
kernel — for the host (launches rewritten to run block by block under the SIMT emulator, a heap-backed stand-in for
the CUDA runtime) and a subset of the GPU parity tests is run against it through WB_LIB, unchanged.

This is NOT a CPU path of the product: nothing under wolkenbase_b200/ builds, ships or selects it, and without a
CUDA device the product library still refuses to create a context (tests/test_abi.py).  It exists so that the
kernels' logic and the ABI's host code are exercised on every CPU-only run — and so that a kernel change can be
checked for parity before a GPU is at hand.  The `-m gpu` run on a B200 remains the proof for the nvcc build."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMT = os.path.join(ROOT, "tests", "simt")
EMULATED = os.path.join(SIMT, "libwolken_b200_emulated.so")

SUBSET = ("pipeline_matches_oracle and 5000 or ragged_and_tiny or return_number_zero or patch_records or "
          "error_behaviour or device_math or street_30k_tile3 or identical_locations or encode_same_layout or "
          "injected_tile")


def test_gpu_parity_subset_passes_on_the_emulated_library():
    subprocess.check_call(["make", "-s", "-C", SIMT, "libwolken_b200_emulated.so"])
    env = dict(os.environ, WB_LIB=EMULATED)
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu",
                          "-q", "-x", "-k", SUBSET, "-p", "no:cacheprovider"], capture_output=True, text=True, env=env,
                         cwd=ROOT, timeout=900)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    assert " passed" in out.stdout and "failed" not in out.stdout, tail


def test_bundled_tile_scan_passes_on_the_emulated_library():
    """The tile scan with 32 and with 5 tiles per warp (small scenes get 1 from wb_scan_bundle): same tile tables."""
    subprocess.check_call(["make", "-s", "-C", SIMT, "libwolken_b200_emulated.so"])
    for bundle in ("32", "5"):
        env = dict(os.environ, WB_LIB=EMULATED, WB_SCAN_BUNDLE=bundle)
        out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu",
                              "-q", "-x", "-k", "pipeline_matches_oracle and 5000 or ragged_and_tiny or street_30k_tile3",
                              "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
        assert out.returncode == 0, bundle + "\n" + out.stdout[-3000:] + out.stderr[-2000:]


def test_sharded_path_passes_on_the_emulated_library():
    """wb_shard_run with ranks as host threads (LOCAL transport): header offsets that differ between ranks, records at
    the XYZ of another rank's point, records dropped by the return-number rule — labels equal the oracle's."""
    subprocess.check_call(["make", "-s", "-C", SIMT, "libwolken_b200_emulated.so"])
    env = dict(os.environ, WB_LIB=EMULATED, WB_EMULATED="1")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_multigpu_gpu.py"), "-m", "gpu",
                          "-q", "-x", "-k", "header_offsets or identical_locations or dropped_records or 1-2-20000", "-p",
                          "no:cacheprovider"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    assert "4 passed" in out.stdout, tail


def test_host_cli_passes_on_the_emulated_library():
    """wolkencli / wolkenquery (the reference-shaped C++ surface over the ABI) with the emulated library preloaded:
    the reference-mode dump, the CLI's own readPoint + embufferPoint flow, and the OctStore queries."""
    subprocess.check_call(["make", "-s", "-C", SIMT, "libwolken_b200_emulated.so"])
    env = dict(os.environ, LD_PRELOAD=EMULATED, WB_EMULATED="1")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_host_cli.py"), "-m", "gpu",
                          "-q", "-x", "-k", "reference_mode or embuffer or octstore_queries or census_after_writing or cli_gpus_equals_oracle_and_one_gpu and 2-3",
                          "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    assert " passed" in out.stdout and "failed" not in out.stdout, tail


def test_results_do_not_depend_on_block_order():
    """The GPU schedules blocks in no particular order; the emulator normally runs them 0..N-1.  Reversed and
    strided orders must give the same dump, tile table and labels (atomically built lists, duplicate handling)."""
    subprocess.check_call(["make", "-s", "-C", SIMT, "libwolken_b200_emulated.so"])
    for order in ("reverse", "shuffle"):
        env = dict(os.environ, WB_LIB=EMULATED, SIMT_BLOCK_ORDER=order)
        out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m",
                              "gpu", "-q", "-x", "-k", "pipeline_matches_oracle and 5000 or ragged_and_tiny or aerial_20k_dups",
                              "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
        assert out.returncode == 0, order + "\n" + out.stdout[-3000:] + out.stderr[-2000:]


def test_round2_candidates_pass_on_the_emulated_library():
    """The classify candidates that are off in the shipped library (WB_CL_XWANTS, WB_CL_REFILTER, WB_CL_COMPACT2),
    all switched on in an emulated build: same dumps, tile tables and labels."""
    subprocess.check_call(["make", "-s", "-C", SIMT, "libwolken_b200_emulated.so"])        # (re)generates gen/
    lib = os.path.join(SIMT, "libwolken_b200_emulated_candidates.so")
    subprocess.check_call(["g++", "-std=gnu++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-Ishim", "-I.",
                           "-Igen", "-DWB_CL_XWANTS=1", "-DWB_CL_REFILTER=1", "-DWB_CL_COMPACT2=1", "-Wno-unused-function",
                           "-o", lib, "gen/wolken_b200.cu.cpp"], cwd=SIMT)
    env = dict(os.environ, WB_LIB=lib)
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu",
                          "-q", "-x", "-k", "pipeline_matches_oracle and (5000 or 20000) or ragged_and_tiny or aerial_20k_dups "
                          "or street_30k_tile3", "-p", "no:cacheprovider"], capture_output=True, text=True, env=env,
                         cwd=ROOT, timeout=900)
    tail = out.stdout[-3000:] + out.stderr[-2000:]
    assert out.returncode == 0, tail
    assert " passed" in out.stdout and "failed" not in out.stdout, tail
