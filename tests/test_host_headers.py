"""LasHeader::openRead on damaged headers (round-1 advice): the CLI must turn them down before it touches the GPU —
these run without one."""
import os
import subprocess

import numpy as np

from wolkenbase_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "wolkenbase_b200", "host", "wolkencli")


def _write(tmp_path, name, mutate):
    cloud = synth.generate(2, 2000, seed=3)
    path = str(tmp_path / name)
    cloud.write(path)
    raw = bytearray(open(path, "rb").read())
    mutate(raw)
    open(path, "wb").write(bytes(raw))
    return path


def test_point_offset_beyond_the_file_leaves_no_points(tmp_path):
    def far(raw):
        raw[96:100] = np.uint32(len(raw) + 12345).tobytes()         # offset to point data
    out = subprocess.run([CLI, "--dump", str(tmp_path / "d"), _write(tmp_path, "far.las", far)],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode != 0 or " 0 points" in out.stdout
    assert out.returncode >= 0, "killed by a signal: %d" % out.returncode


def test_point_count_that_would_wrap_is_clamped(tmp_path):
    def huge(raw):
        raw[107:111] = np.uint32(0xffffffff).tobytes()              # legacy number of point records
    path = _write(tmp_path, "huge.las", huge)
    out = subprocess.run([CLI, "--dump", str(tmp_path / "d"), path], capture_output=True, text=True, timeout=120)
    assert out.returncode >= 0, "killed by a signal: %d" % out.returncode
    assert "4294967295 points" not in out.stdout


def test_record_length_below_the_formats_minimum_is_not_a_las_file(tmp_path):
    def short(raw):
        raw[105:107] = np.uint16(12).tobytes()                      # point data record length (format 1 needs 28)
    out = subprocess.run([CLI, "--dump", str(tmp_path / "d"), _write(tmp_path, "short.las", short)],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 1 and "not a LAS file" in out.stderr


def test_cloud_surface_and_block_census_without_a_gpu():
    """cloud.cpp's block view (getNumCloudBlocks / getCloudBlock) and testpattern.cpp's per-block census, in the
    reference-shaped host library: wolkenbase_b200/host/hosttest.cpp (assertions, as the reference's wolkentest)."""
    exe = os.path.join(ROOT, "wolkenbase_b200", "host", "hosttest")
    subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "hosttest ok" in out.stdout, out.stdout + out.stderr
