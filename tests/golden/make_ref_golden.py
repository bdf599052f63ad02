"""Generates tests/golden/ref_*.json: the stdout "Output" block of the REFERENCE ITSELF (oracle/_ref/montecarlo_ref, built by
`make -C oracle ref` in the dev container from the sources under /root/reference) for the cases in tests/refcases.py, with
MCREF_SEED=0 and one thread.  Run from the repo root:   python tests/golden/make_ref_golden.py
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbin                      # noqa: E402
from tests import refcases                     # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    allc = dict(refcases.CASES); allc.update(refcases.REF_ONLY)
    for name, case in allc.items():
        d = tempfile.mkdtemp(prefix="mcref_")
        mat, _ = refcases.write_material(case, d)
        dom, prob = refcases.ref_argv(case)
        blocks, out, dt = refbin.run(d, mat, case["T"], dom, prob, seed=0, threads=1, checked=True)
        esc = [ln for ln in out.splitlines() if "esc:" in ln][-1].split("esc:")[1].strip()
        rec = {"case": name, "argv": [mat, case["T"]] + dom + prob, "seed": 0, "threads": 1, "esc": int(esc),
               "output": blocks["Output"][0].tolist(),
               "averaged": blocks["Averaged"][0].tolist() if "Averaged" in blocks else None,
               "how": "oracle/_ref/montecarlo_ref_chk (all assertions on), MCREF_SEED=0, OMP_NUM_THREADS=1"}
        json.dump(rec, open(os.path.join(HERE, f"ref_{name}.json"), "w"), indent=0)
        print(f"{name}: {len(rec['output'])}x{len(rec['output'][0])}  esc {esc}  {dt:.1f} s")


if __name__ == "__main__":
    main()
