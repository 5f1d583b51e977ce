"""Helpers shared by the golden-vector tests (oracle and GPU engine read the same fixtures)."""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
MIN_SCORE = -858993459
_CODE = {"M": 0, "S": 1, "D": 2, "I": 3}


def load_cases(name="pairwise_vectors.json"):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)["cases"]


def parse_ops(s):
    """'Y4 5MS3M' -> [(5,4),(0,0)x5,(1,0),(0,0)x3] as (code, clip_len)."""
    out = []
    for chunk in s.split():
        if chunk[0] in "XY":
            out.append((4 if chunk[0] == "X" else 5, int(chunk[1:])))
            continue
        for cnt, letter in re.findall(r"(\d*)([MSDI])", chunk):
            out.extend([(_CODE[letter], 0)] * (int(cnt) if cnt else 1))
    return out


def clip(v):
    return MIN_SCORE if v == "MIN" else int(v)


def scoring_fields(sc):
    """Golden 'scoring' dict -> kwargs common to oracle.make_scoring and rust_bio_b200 Scoring."""
    return dict(
        gap_open=sc["gap_open"], gap_extend=sc["gap_extend"],
        match=sc.get("match", 0), mismatch=sc.get("mismatch", 0), matrix=sc.get("matrix"),
        from_scores=bool(sc.get("from_scores", False)),
        xclip_prefix=clip(sc.get("xclip_prefix", "MIN")), xclip_suffix=clip(sc.get("xclip_suffix", "MIN")),
        yclip_prefix=clip(sc.get("yclip_prefix", "MIN")), yclip_suffix=clip(sc.get("yclip_suffix", "MIN")))
