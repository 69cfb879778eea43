"""
Fusion of per-robot odometry g2o files and loop-closure g2o files (as written by roman_b200.align.results.write_g2o)
into the single pose-graph file the reference hands to Kimera-RPGO — counterpart of the reference's
roman/offline_rpgo/g2o_file_fusion.py:8-142.

Vertex names are rewritten to GTSAM symbol keys: key = (ord(letter) << 56) | index (gtsam::Symbol; the reference
calls gtsam.symbol(letter, index)).  A loop-closure edge is kept only if the `# LC: <n>` comment right above it
reports at least `thresh` associations.  For single-robot files an edge is kept once: when its first vertex token
sorts strictly before its second — the reference compares the raw TOKENS (strings), and so does this.
"""
from __future__ import annotations

import os
from typing import Iterator, List, Optional, Tuple


def gtsam_symbol(letter: str, index: int) -> int:
    return (ord(letter) << 56) | int(index)


def _digits(token: str) -> int:
    return int("".join(ch for ch in token if ch.isdigit()))


def _records(path: str) -> Iterator[Tuple[List[str], str]]:
    """(tokens, previous raw line) for every VERTEX_SE3:QUAT / EDGE_SE3:QUAT line of a g2o file."""
    with open(os.path.expandvars(path), "r") as f:
        raw = f.readlines()
    for k, text in enumerate(raw):
        text = text.strip()
        if not text or text.startswith("#"):
            continue
        tok = text.split()
        ok = (tok[0] == "EDGE_SE3:QUAT" and len(tok) == 31) or (tok[0] == "VERTEX_SE3:QUAT" and len(tok) == 9)
        assert ok, f"Invalid line: {tok}"
        yield tok, (raw[k - 1].strip() if k > 0 else "")


_EDGE_GROUPS = (3, 4, 6, 5, 4, 3, 2, 1)   # translation, quaternion, then the rows of the upper-triangular information matrix


def _edge_text(tok: List) -> str:
    out = f"{tok[0]}\t{tok[1]} {tok[2]} \t"
    pos = 3
    for g in _EDGE_GROUPS[:-1]:
        out += " ".join(str(v) for v in tok[pos:pos + g]) + " \t"
        pos += g
    return out + str(tok[pos])


def vertex_lines(path: str, letter: str) -> List[str]:
    out = []
    for tok, _ in _records(path):
        if tok[0] != "VERTEX_SE3:QUAT":
            continue
        out.append(" ".join([tok[0], str(gtsam_symbol(letter, _digits(tok[1])))] + tok[2:]))
    return out


def edge_lines(path: str, letter1: str, letter2: str, thresh: Optional[int] = None, lc: bool = False,
               self_lc: bool = False) -> List[str]:
    out = []
    for tok, prev in _records(path):
        if tok[0] != "EDGE_SE3:QUAT":
            continue
        if self_lc and tok[1] >= tok[2]:
            continue
        if lc:
            assert prev.startswith("# LC:"), "loop closure must be preceded by a comment"
            if thresh is not None and int(prev.split()[2]) < thresh:
                continue
        tok = list(tok)
        tok[1], tok[2] = gtsam_symbol(letter1, _digits(tok[1])), gtsam_symbol(letter2, _digits(tok[2]))
        out.append(_edge_text(tok))
    return out


def create_config(robots, odometry_g2o_dir, submap_align_dir=None, align_file_name=None) -> dict:
    """Robot letters a, b, c, ... in list order; odometry `<dir>/<robot>.g2o`; loop closures
    `<align dir>/<r1>_<r2>/<name>.g2o` for r1 == r2 (single) and every r1 before r2 (multi)."""
    cfg = {"robots": [], "odometry": [], "single_lc": [], "multi_lc": []}
    for i, r in enumerate(robots):
        cfg["robots"].append({"robot": r, "letter": chr(ord("a") + i)})
        cfg["odometry"].append({"robot": r, "file": f"{odometry_g2o_dir}/{r}.g2o"})
        if submap_align_dir is None:
            continue
        cfg["single_lc"].append({"robot": r, "file": f"{submap_align_dir}/{r}_{r}/{align_file_name}.g2o"})
        for r2 in robots[i + 1:]:
            cfg["multi_lc"].append({"robot1": r, "robot2": r2, "file": f"{submap_align_dir}/{r}_{r2}/{align_file_name}.g2o"})
    return cfg


def g2o_file_fusion(config: dict, output: str, thresh: Optional[int] = None):
    letter = {r["robot"]: r["letter"] for r in config["robots"]}
    lines: List[str] = []
    for od in config["odometry"]:
        a = letter[od["robot"]]
        lines += edge_lines(od["file"], a, a, thresh, lc=False)
        lines += vertex_lines(od["file"], a)
    for sl in config["single_lc"]:
        a = letter[sl["robot"]]
        lines += edge_lines(sl["file"], a, a, thresh, lc=True, self_lc=True)
    for ml in config["multi_lc"]:
        lines += edge_lines(ml["file"], letter[ml["robot1"]], letter[ml["robot2"]], thresh, lc=True)
    with open(output, "w") as f:
        for ln in lines:
            f.write(ln + "\n")
