"""nhwcodec_b200/build.py -- compile libnhw_cuda.so for sm_100a, in-tree.

    python -m nhwcodec_b200.build          (or __graft_entry__.build())

nvcc cross-compiles without a GPU.  The shared object is git-ignored but travels to the GPU
box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnhw_cuda.so")
SOURCES = ["api.cu", "front.cu", "front_fused.cu", "synth.cu", "encode.cu", "decode.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # integer/IEEE-exact pipeline: never contract a*b+c
    "--extended-lambda",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-cudart", "static",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "nhw_cuda.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(CSRC, s[:-3] + ".o")
        cmd = ["nvcc", *NVCC_FLAGS, "-Xcompiler", "-fvisibility=default", "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = ["nvcc", "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    subprocess.check_call(cmd)
    build_host()
    return LIB


def build_host():
    """libnhw_compat.so (the reference's per-image entry points) and the nhw-enc CLI: plain C."""
    compat = os.path.join(HERE, "libnhw_compat.so")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-Wall", os.path.join(CSRC, "compat.c"), "-o", compat,
                           "-L" + HERE, "-lnhw_cuda", "-Wl,-rpath,$ORIGIN"])
    cli = os.path.join(HERE, "..", "cli")
    subprocess.check_call(["gcc", "-O2", "-Wall", os.path.join(cli, "nhw_enc_cli.c"), "-o", os.path.join(cli, "nhw-enc"),
                           "-L" + HERE, "-lnhw_compat", "-lnhw_cuda", "-Wl,-rpath,$ORIGIN/../nhwcodec_b200"])
    compat_dec = os.path.join(HERE, "libnhw_compat_dec.so")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-Wall", os.path.join(CSRC, "compat_dec.c"), "-o", compat_dec,
                           "-L" + HERE, "-lnhw_cuda", "-Wl,-rpath,$ORIGIN"])
    subprocess.check_call(["gcc", "-O2", "-Wall", os.path.join(cli, "nhw_dec_cli.c"), "-o", os.path.join(cli, "nhw-dec"),
                           "-L" + HERE, "-lnhw_cuda", "-Wl,-rpath,$ORIGIN/../nhwcodec_b200"])
    # batch I/O front-end (image readers, tiling, the .nhwpack container, manifest jobs) and its CLI
    batchio = os.path.join(HERE, "libnhw_batchio.so")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-fvisibility=hidden", os.path.join(CSRC, "batchio.c"), "-o", batchio,
                           "-L" + HERE, "-lnhw_cuda", "-lpthread", "-Wl,-rpath,$ORIGIN"])
    subprocess.check_call(["gcc", "-O2", "-Wall", os.path.join(cli, "nhw_batch_cli.c"), "-o", os.path.join(cli, "nhw-batch"),
                           "-L" + HERE, "-lnhw_batchio", "-lnhw_cuda", "-Wl,-rpath,$ORIGIN/../nhwcodec_b200"])
    # drop-in proof: the reference's OWN, unmodified CLI source compiled against its own header
    # and linked against our two libraries (only where the reference tree is present)
    ref_cli = "/root/reference/encoder/nhw_encoder_cli.c"
    ref_out = os.path.join(HERE, "..", "oracle", "_ref")
    if os.path.exists(ref_cli) and os.path.isdir(ref_out):
        subprocess.check_call(["gcc", "-O2", "-w", "-I/root/reference/encoder", ref_cli, "-o",
                               os.path.join(ref_out, "nhw-enc-dropin"), "-L" + HERE, "-lnhw_compat", "-lnhw_cuda",
                               "-Wl,-rpath,$ORIGIN/../../nhwcodec_b200"])
        subprocess.check_call(["gcc", "-O2", "-w", "-ffp-contract=off", "-I/root/reference/decoder",
                               "/root/reference/decoder/nhw_decoder_cli.c", "-o", os.path.join(ref_out, "nhw-dec-dropin"),
                               "-L" + HERE, "-lnhw_compat_dec", "-lnhw_cuda", "-Wl,-rpath,$ORIGIN/../../nhwcodec_b200"])


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
