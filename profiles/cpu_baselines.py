#!/usr/bin/env python
"""profiles/cpu_baselines.py -- the two CPU-side reports BASELINE.md section 3 asks for next to the GPU numbers, run on the
GPU box's host cores with the reference compiled there by oracle/build_ref.sh (oracle/_ref, test infrastructure):

  cli_throughput : what a user of the reference gets -- the STOCK CLIs (README one-liner build, gcc -O3) over a batch of BMP
                   files in tmpfs with `xargs -P $(nproc)`, one process per image, encode then decode.
  stock_match    : per image, whether the stock encoder's .nhw equals the canonical (zero-guard) encoder's outside the
                   don't-care bits of SURVEY.md Appendix A (the last byte of select_word1/2 and of res1/3/5/6_word), and
                   whether both files decode to the same pixels ("k of n").

One JSON object on stdout.  Nothing here is a product path."""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")
DONT_CARE = ("select_word1", "select_word2", "res1_word", "res3_word", "res5_word", "res6_word")


def bmp_header():
    import struct
    return b"BM" + struct.pack("<IHHI", 54 + 786432, 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, 512, 512, 1, 24, 0, 786432, 2835, 2835, 0, 0)


def masked_equal(a, b):
    """equal outside the don't-care bytes?  -> (bool, first differing section or None)"""
    from nhwcodec_b200 import container
    ha, sa = container.parse_nhw(a)
    hb, sb = container.parse_nhw(b)
    for k in ha:
        if k not in ("parsed_bytes", "file_bytes") and ha[k] != hb.get(k):
            return False, "header:" + k
    for k in sa:
        x, y = sa[k], sb.get(k, b"")
        if k in DONT_CARE and len(x) == len(y):
            x, y = x[:-1], y[:-1]
        if x != y:
            return False, k
    return True, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=256, help="files in the CLI-throughput batch")
    ap.add_argument("--match-images", type=int, default=48)
    ap.add_argument("--quality", type=int, default=20)
    args = ap.parse_args()
    from nhwcodec_b200 import synth
    cores = len(os.sched_getaffinity(0))
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="nhwcli_", dir=base)
    out = {"cores": cores, "cpu": next((l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "?")}
    try:
        hd = bmp_header()
        distinct = min(args.images, 32)
        pix = [synth.natural(1000 + i) for i in range(distinct)]
        for i in range(args.images):
            with open(os.path.join(d, "i%05d.bmp" % i), "wb") as f:
                f.write(hd)
                f.write(pix[i % distinct].tobytes())
        # ---- (i) CLI throughput, stock build
        enc, dec = os.path.join(REF, "nhw-enc-stock"), os.path.join(REF, "nhw-dec-stock")
        names = "\n".join("i%05d" % i for i in range(args.images)).encode()
        t0 = time.perf_counter()
        subprocess.run(["xargs", "-P", str(cores), "-I", "{}", enc, "-q%d" % args.quality, "{}.bmp", "{}.nhw"], input=names, cwd=d,
                       check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t_enc = time.perf_counter() - t0
        t0 = time.perf_counter()
        subprocess.run(["xargs", "-P", str(cores), "-I", "{}", dec, "{}.nhw", "{}.out.bmp"], input=names, cwd=d, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        t_dec = time.perf_counter() - t0
        mpix = args.images * 0.262144
        out["cli_throughput"] = {
            "what": "stock reference CLIs (gcc -O3), %d BMP files in %s, xargs -P %d, one process per image, -q%d" % (args.images, base or "tmp", cores, args.quality),
            "encode_MPix_s": round(mpix / t_enc, 1), "decode_MPix_s": round(mpix / t_dec, 1),
            "round_trip_MPix_s": round(mpix / (t_enc + t_dec), 1), "encode_s": round(t_enc, 3), "decode_s": round(t_dec, 3)}
        # ---- (ii) stock vs canonical, per image
        canon, dcanon = os.path.join(REF, "nhw-enc-canon"), os.path.join(REF, "nhw-dec-canon")
        gens = [("natural-like", synth.natural, 2000), ("textured", synth.textured, 2100), ("uniform-noise", synth.noise, 2200)]
        report = []
        for q in sorted({args.quality, 12, 23}):
            for gname, gen, seed0 in gens:
                n = args.match_images if gname == "natural-like" else max(args.match_images // 6, 4)
                same = masked = decodes = 0
                where = {}
                for i in range(n):
                    p = os.path.join(d, "m.bmp")
                    with open(p, "wb") as f:
                        f.write(hd)
                        f.write(gen(seed0 + i).tobytes())
                    subprocess.run([enc, "-f", "-q%d" % q, p, os.path.join(d, "m_stock.nhw")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                    subprocess.run([canon, "-f", "-q%d" % q, p, os.path.join(d, "m_canon.nhw")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                    a, b = open(os.path.join(d, "m_stock.nhw"), "rb").read(), open(os.path.join(d, "m_canon.nhw"), "rb").read()
                    same += a == b
                    ok, sec = masked_equal(a, b) if len(a) and len(b) else (False, "empty")
                    masked += ok
                    if not ok:
                        where[sec] = where.get(sec, 0) + 1
                    subprocess.run([dec, os.path.join(d, "m_stock.nhw"), os.path.join(d, "m_s.bmp")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                    subprocess.run([dcanon, os.path.join(d, "m_canon.nhw"), os.path.join(d, "m_c.bmp")], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                    decodes += hashlib.md5(open(os.path.join(d, "m_s.bmp"), "rb").read()).digest() == hashlib.md5(open(os.path.join(d, "m_c.bmp"), "rb").read()).digest()
                report.append({"quality": q, "generator": gname, "n": n, "byte_identical": same, "match_outside_dont_care_bits": masked,
                               "decode_to_identical_pixels": decodes, "first_differing_section_counts": where})
        out["stock_match"] = {"what": "stock glibc build vs canonical (zero-guard) build of the unmodified reference encoder, one process per image; "
                                      "don't-care bits = last byte of " + ", ".join(DONT_CARE) + " (SURVEY.md Appendix A)", "cases": report}
    finally:
        shutil.rmtree(d, ignore_errors=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
