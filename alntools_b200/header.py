"""BAM header -> target / haplotype tables (alntools/bam_utils.py:561-633, multisample :399-465).

Stays on the host: O(#@SQ) string work done once per job.  The device only ever sees the resulting
int32 lookups tid -> (main-target index, sorted-haplotype index).
"""
from collections import OrderedDict
import sys

import numpy as np

from . import utils

LOG = utils.get_logger()


class TargetTables(object):
    """main_targets (OrderedDict name -> idx), haplotypes (sorted list), lengths int32[T, H],
    tid_target / tid_hap int32[n_references]."""

    def __init__(self, references, ref_lengths, target_filename=None):
        main_targets = OrderedDict()
        if target_filename:                                   # bam_utils.py:571-579
            main_targets = utils.parse_targets(target_filename)
            if len(main_targets) == 0:
                LOG.error("Unable to parse target file")
                sys.exit(-1)
        pieces = []
        haplotypes = set()
        for name in references:                               # :582-600
            i = name.rfind("_")
            target, hap = (name[:i], name[i + 1:]) if i > 0 else (name, "")
            pieces.append((target, hap))
            haplotypes.add(hap)
            if target not in main_targets:
                main_targets[target] = len(main_targets)
        self.main_targets = main_targets
        self.haplotypes = sorted(haplotypes)                  # :602 ('' sorts first)
        hap_idx = {h: i for i, h in enumerate(self.haplotypes)}
        n = len(references)
        self.lengths = np.zeros((len(main_targets), len(self.haplotypes)), dtype=np.int32)  # :605
        self.tid_target = np.empty(n, dtype=np.int32)
        self.tid_hap = np.empty(n, dtype=np.int32)
        for tid, (target, hap) in enumerate(pieces):          # :615-633
            t, h = main_targets[target], hap_idx[hap]
            self.lengths[t, h] = ref_lengths[tid]
            self.tid_target[tid] = t
            self.tid_hap[tid] = h
        # The reference rebuilds '<target>_<hap>' and looks the name up again (bam_utils.py:800-811);
        # that only differs from the direct tid -> (target, hap) map when two @SQ names collapse to the
        # same pair (e.g. 'a' and 'a_'), which no aligner index produces.  Refuse instead of guessing.
        seen = {}
        for tid, pair in enumerate(pieces):
            if pair in seen:
                raise ValueError("@SQ names %r and %r map to the same (target, haplotype)"
                                 % (references[seen[pair]], references[tid]))
            seen[pair] = tid

    @property
    def num_targets(self):
        return len(self.main_targets)

    @property
    def num_haplotypes(self):
        return len(self.haplotypes)
