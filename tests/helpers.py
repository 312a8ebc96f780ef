import json
import os
import re

import numpy as np

FIELD_ORDER = ["max", "zdropped", "max_q", "max_t", "mqe", "mqe_t", "mte", "mte_q", "score", "n_cigar"]


def load_json(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


def parse_cigar(s: str, ops="MID"):
    return [(int(n) << 4) | ops.index(o) for n, o in re.findall(r"(\d+)([A-Z])", s)]


def cigar_consistent(cigar, qlen, tlen, fields, flag):
    """Size-independent invariants of a ksw_extz2 CIGAR (forward order): M+I consume the query prefix
    ending at the traceback start, M+D the target prefix."""
    q = sum(c >> 4 for c in cigar if (c & 0xF) in (0, 1))
    t = sum(c >> 4 for c in cigar if (c & 0xF) in (0, 2))
    if not fields["zdropped"] and not (flag & 0x40):
        return q == qlen and t == tlen
    if fields["max_t"] >= 0 and fields["max_q"] >= 0:
        return q == fields["max_q"] + 1 and t == fields["max_t"] + 1
    return len(cigar) == 0
