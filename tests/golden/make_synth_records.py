"""Records what the UNMODIFIED reference (oracle/_ref/libslip_ref.so) computes for synthetic systems
of the BASELINE config families at sizes beyond the explicit fixtures: digests of L, U, rhos, pinv
and x, sizes, and the reference's seconds in the build container -> tests/golden/synth_records.json.

Run in the build container:   python tests/golden/make_synth_records.py [NAME ...]
The systems themselves are not stored: slip_lu_b200.refmats.synth_system(name) regenerates them
from their seeds (same generator here and on the GPU box).  Minutes of CPU per record.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from slip_lu_b200 import capi, refmats  # noqa: E402
from oracle import binding as ob        # noqa: E402
from make_refmats import reference_record  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synth_records.json")


def main():
    ob.build()
    ref = capi.SlipLib(ob.REF_SO)
    names = sys.argv[1:] or list(refmats.SYNTH)
    try:
        with open(OUT) as f:
            records = {r["name"]: r for r in json.load(f)["records"]}
    except FileNotFoundError:
        records = {}
    for name in names:
        n, I, J, X, b = refmats.synth_system(name)
        records[name] = reference_record(ref, name, n, I, J, X, b, refmats.SYNTH[name]["family"])
        with open(OUT, "w") as f:
            json.dump(dict(note="outputs of the unmodified reference (default options) on synthetic systems "
                                "regenerated from their seeds by slip_lu_b200.refmats.synth_system; "
                                "ref_seconds measured in the build container",
                           records=[records[k] for k in sorted(records)]), f, indent=0)


if __name__ == "__main__":
    main()
