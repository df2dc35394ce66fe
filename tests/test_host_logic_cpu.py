"""CPU checks of closed forms / tables the CUDA kernels use in place of the reference's branchy rules
(tests/hostcheck/check_tables.cpp compiles the shared headers with g++ and compares both forms)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tables_match_branchy_rules(tmp_path):
    exe = str(tmp_path / "check_tables")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "hostcheck", "check_tables.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "sround mismatches 0" in out.stdout and "pair table mismatches 0" in out.stdout


def test_codebook_error_exits(tmp_path):
    """NHW_ERR_CODEBOOK stands for the reference's exit(-1) (encoder/compress_pixel.c:234,270-271); no input we know reaches it
    on the GPU, so the branches of the shared __host__ __device__ functions are driven with hand-made histograms"""
    exe = str(tmp_path / "check_codebook")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "hostcheck", "check_codebook.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "codebook checks failed: 0" in out.stdout, out.stdout
