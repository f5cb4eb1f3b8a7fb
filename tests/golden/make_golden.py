"""Generates tests/golden/*.npz by running the REFERENCE's own code (oracle/_ref/libswtpg_ref.so, i.e. the unmodified
headers under /root/reference compiled by oracle/Makefile) on deterministic inputs. Run it in the build container:

    make -C oracle ref && python tests/golden/make_golden.py

The fixtures hold only the reference OUTPUTS (TP tuples, final pedestals); the inputs are re-created at test time by
`tests/cases.py` from the seeds recorded here, so nothing from /root/reference is needed when the tests run.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from fdreadoutlibs_b200 import frames as F  # noqa: E402
from oracle import binding as B  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    assert B.reference_available(), "build oracle/_ref first (make -C oracle ref)"
    blobs = {}
    for name, case in cases.GOLDEN_CASES.items():
        units = cases.make_input(case)  # [n_links, n_units, unit_bytes]
        tps_all, peds = [], []
        for l in range(units.shape[0]):
            if case["fmt"] == "wibeth":
                r = B.ReferenceWibEth(case["ref_impl"], case["threshold"], case.get("acc_limit", 10), case.get("rs_memory_factor", 8),
                                      case.get("rs_scale_factor", 5), link_id=l)
                tps, ped = r.process(units[l], dump=True)
                peds.append(ped[-1, 0])
            else:
                r = B.ReferenceWib2(case["ref_impl"], case["threshold"], link_id=l)
                tps, st = r.process(units[l], dump=True)
                peds.append(st[-1, 0])
            tps_all.append(tps)
        tps = F.sort_tps(np.concatenate(tps_all))
        blobs[name + "__tps"] = tps
        blobs[name + "__pedestal"] = np.stack(peds)
        print(f"{name}: {tps.size} TPs")
    # unpack known-answer vectors straight from the reference's expansion functions
    lib = B.ref_lib()
    fr = cases.unpack_kat_frame()
    out = np.zeros(4096, dtype=np.uint16)
    lib.ref_wibeth_expand(fr.ctypes.data, out.ctypes.data)
    blobs["unpack_kat_wibeth"] = out
    sc = cases.unpack_kat_superchunk()
    for sel in (0, 1):
        o2 = np.zeros(8 * 12 * 16, dtype=np.uint16)
        lib.ref_wib2_expand(sc.ctypes.data, sel, o2.ctypes.data)
        blobs[f"unpack_kat_wib2_sel{sel}"] = o2
    taps = np.zeros(7, dtype=np.int16)
    lib.ref_firwin_int(7, 0.1, 64, taps.ctypes.data)
    blobs["firwin_int_7_0p1_64"] = taps
    np.savez_compressed(os.path.join(OUT, "reference_vectors.npz"), **blobs)
    print("wrote", os.path.join(OUT, "reference_vectors.npz"), os.path.getsize(os.path.join(OUT, "reference_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
