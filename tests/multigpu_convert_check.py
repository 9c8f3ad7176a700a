"""Multi-GPU convert() on a box with >= 2 GPUs: every golden single-sample case and one larger synthetic BAM
through `python -m alntools_b200 bam2ec` with ALNTOOLS_GPUS = 2 .. all GPUs; the EC file must be the golden
(or the single-GPU) file byte for byte.  usage: python tests/multigpu_convert_check.py [max_gpus]"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def run_cli(bam, out, gpus, extra=()):
    env = dict(os.environ, ALNTOOLS_GPUS=str(gpus), PYTHONPATH=ROOT)
    subprocess.check_call([sys.executable, "-m", "alntools_b200", "bam2ec", bam, out] + list(extra), env=env, cwd=ROOT)


def main():
    import torch
    n_gpus = torch.cuda.device_count()
    if len(sys.argv) > 1:
        n_gpus = min(n_gpus, int(sys.argv[1]))
    assert n_gpus >= 2, "needs at least two GPUs"
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        cases = [c for c in json.load(fh) if c["kind"] == "single"]
    worlds = sorted(set([2, n_gpus]))
    with tempfile.TemporaryDirectory() as tmp:
        for case in cases:
            extra = ["-t", os.path.join(GOLDEN, case["targets"])] if case.get("targets") else []
            for w in worlds:
                out = os.path.join(tmp, "%s.%d.bin" % (case["name"], w))
                run_cli(os.path.join(GOLDEN, case["bam"]), out, w, extra)
                with open(out, "rb") as a, open(os.path.join(GOLDEN, case["ec"]), "rb") as b:
                    assert a.read() == b.read(), (case["name"], w)
                print("golden %s on %d GPUs: identical" % (case["name"], w), flush=True)
        # a file large enough that every rank has real work: against the single-GPU file
        from alntools_b200 import synth
        cols = synth.make_columns(1_500_000, 20000, 2, seed=77, mode="diploid", dup_rate=0.01)
        bam = os.path.join(tmp, "big.bam")
        synth.columns_to_bam(bam, cols, 20000, 2)
        one = os.path.join(tmp, "big.1.bin")
        run_cli(bam, one, 1)
        for w in worlds:
            out = os.path.join(tmp, "big.%d.bin" % w)
            rng = os.path.join(tmp, "big.%d.range" % w)
            run_cli(bam, out, w, ["--rangefile", rng])
            with open(out, "rb") as a, open(one, "rb") as b:
                assert a.read() == b.read(), ("big", w)
            print("synthetic 1.5 M reads on %d GPUs: identical to one GPU (%d bytes)" % (w, os.path.getsize(out)), flush=True)
    print("multigpu convert check: ok")


if __name__ == "__main__":
    main()
