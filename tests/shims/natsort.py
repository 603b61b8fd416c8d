"""Stand-in for `natsort` (absent offline)."""
import re


def natsorted(seq):
    return sorted(seq, key=lambda s: [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", str(s))])
