#!/usr/bin/env python
"""Provenance of tests/golden/reference_known_answers.json.

The fixture holds known answers that the reference (Knockoffs.jl v2.0.2) prints in its own docs, notebooks and tests;
Julia is not installed in the build image, so they are TRANSCRIBED, not generated.  This script re-reads the cited
source lines from a checkout of the reference and checks that every number of every fixture entry occurs there, in
order (and, for docs tables, in the stated column).  It only needs the reference tree at build time:

    python tests/golden/make_golden.py [/root/reference]        # exit 0 = the fixture matches its sources

Nothing in tests/ or bench.py reads the reference tree at run time (the GPU box does not have it).
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
FLOAT = re.compile(r"(?<![\w.])-?\d+\.\d+(?:[eE][-+]?\d+)?|(?<![\w.])-?\d+(?![\w.])")


def lines(ref, path, lo, hi):
    with open(os.path.join(ref, path), encoding="utf-8") as f:
        return f.read().splitlines()[lo - 1:hi]


def column(ls, col):
    out = []
    for l in ls:
        nums = FLOAT.findall(l)
        if len(nums) > col and "⋮" not in l:
            out.append(float(nums[col]))
    return out


def is_subsequence(needle, hay, tol=0.0):
    it = iter(hay)
    return all(any(abs(x - y) <= tol for y in it) for x in needle)


def check_vector(name, g, col_values):
    seq = list(g["head"]) + list(g["tail"])
    ok = is_subsequence(seq, col_values)
    print(f"{'ok ' if ok else 'BAD'} {name}: {len(seq)} numbers against {g['source']}")
    return ok


def notebook_first_column(ref, path, cell_index):
    nb = json.load(open(os.path.join(ref, path), encoding="utf-8"))
    code = [c for c in nb["cells"] if c["cell_type"] == "code"]
    vals = []
    for out in code[cell_index].get("outputs", []):
        text = out.get("text") or out.get("data", {}).get("text/plain") or []
        for l in (text if isinstance(text, list) else text.splitlines()):
            nums = FLOAT.findall(l)
            if nums and "⋮" not in l and "Matrix" not in l and "Vector" not in l:
                vals.append(float(nums[0]))
    return vals


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    fx = json.load(open(os.path.join(HERE, "reference_known_answers.json")))
    ok = True
    doc = "docs/src/man/modelX/modelX.md"
    ok &= check_vector("K1", fx["K1_mvr_ar1_rho0.4_p1000_m1"], column(lines(ref, doc, 141, 169), 0))
    tab = lines(ref, doc, 180, 208)
    ok &= check_vector("K2 maxent", fx["K2_maxent_ar1_rho0.4_p1000_m1"], column(tab, 0))
    ok &= check_vector("K2 equi", fx["K2_equi_ar1_rho0.4_p1000_m1"], column(tab, 2))
    ok &= check_vector("K3", fx["K3_maxent_ar1_rho0.4_p1000_m5"], column(lines(ref, doc, 278, 306), 3))
    nbp = "test/compare_knockpy.ipynb"
    nb = json.load(open(os.path.join(ref, nbp), encoding="utf-8"))
    allnums = []
    for c in nb["cells"]:
        if c["cell_type"] != "code":
            continue
        for out in c.get("outputs", []):
            text = out.get("text") or out.get("data", {}).get("text/plain") or []
            for l in (text if isinstance(text, list) else text.splitlines()):
                nums = FLOAT.findall(l)
                if nums and "⋮" not in l:
                    allnums.append(float(nums[0]))
    for key in ("K4_sdp_ccd_ar1_rho0.5_p300_m1", "K4_mvr_ar1_rho0.5_p300_m1", "K4_maxent_ar1_rho0.5_p300_m1"):
        ok &= check_vector(key, fx[key], allnums)
    rt = "\n".join(lines(ref, "test/runtests.jl", 157, 166))
    nums = [float(x) for x in FLOAT.findall(rt)]
    for c in fx["K7_threshold_runtests_157_166"]["cases"]:
        good = is_subsequence(c["w"], nums) and (c["tau"] == float("inf") and "Inf" in rt or is_subsequence([c["tau"]], nums))
        print(f"{'ok ' if good else 'BAD'} K7 threshold q={c['q']} {c['method']}")
        ok &= good
    nums = [float(x) for x in FLOAT.findall("\n".join(lines(ref, "test/runtests.jl", 200, 228)))]
    g = fx["K8_mk_statistics_runtests_200_228"]
    seqs = [g["single"]["beta"], g["single"]["betako"], g["multi"]["T0"]] + g["multi"]["Tk"] + [g["multi"]["tau"], [float(k) for k in g["multi"]["kappa"]]]
    good = all(is_subsequence(s, nums) for s in seqs)
    print(f"{'ok ' if good else 'BAD'} K8 MK_statistics")
    ok &= good
    nums = [float(x) for x in FLOAT.findall("\n".join(lines(ref, "test/runtests.jl", 733, 745)))]
    g = fx["K9_choose_group_reps_issue71_runtests_735_744"]
    flat = [x for row in g["Sigma"] for x in row]
    good = is_subsequence(flat, nums) and is_subsequence([float(x) for x in g["groups"]], nums)
    for c in g["cases"]:
        good = good and is_subsequence([c["threshold"]] + [float(r) for r in c["group_reps"]], nums)
    print(f"{'ok ' if good else 'BAD'} K9 choose_group_reps")
    ok &= good
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
