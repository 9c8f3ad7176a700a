import cProfile, pstats, os, sys, tempfile, time
sys.path.insert(0, os.getcwd())
from alntools_b200 import bam_utils, synth
import torch
torch.cuda.init()
for reads in (8_000_000,):
    cols = synth.make_columns(reads, 100000, 2, 1002, mode="diploid")
    with tempfile.TemporaryDirectory() as tmp:
        bam = os.path.join(tmp, "s.bam")
        synth.columns_to_bam(bam, cols, 100000, 2)
        print("bam MB", os.path.getsize(bam) / 1e6)
        bam_utils.convert(bam, os.path.join(tmp, "o.bin"), None)
        cProfile.run('bam_utils.convert(bam, os.path.join(tmp, "o.bin"), None)', "/tmp/c.prof")
        pstats.Stats("/tmp/c.prof").sort_stats("tottime").print_stats(16)
        print("ec MB", os.path.getsize(os.path.join(tmp, "o.bin")) / 1e6)
