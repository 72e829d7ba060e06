"""Regenerate tests/golden/ from the reference checkout (run in the build container only).

Two kinds of fixtures:

1. Verbatim copies of the reference's own *test data* (inputs and golden outputs, not source
   code) from ``/root/reference/tests/data`` -- the GPU box has no reference checkout.
2. ``sugar_params.json``: equilibrium values / force constants of the sugar test case computed by
   the oracle restatement of the *current* reference algorithm (circular statistics), after the
   oracle has been checked against ``out.rtp``.  ``sugar_out.itp`` in the reference is stale
   (arithmetic statistics, see SURVEY.md "Five facts" no. 3), so it cannot serve at 1e-4.

Usage:  python tests/golden/make_golden.py
"""

import json
import pathlib
import shutil
import sys

HERE = pathlib.Path(__file__).resolve().parent
REF = pathlib.Path("/root/reference/tests/data")

COPIES = [
    "sugar.gro", "sugar.xtc", "sugar.map", "sugar.bnd", "sugar_out.xtc", "sugar-cg.gro", "sugar_out.itp",
    "sugar_out.gro", "sugar_only.map",
    "ALLA_length.dat", "ALLA_angle.dat", "ALLA_dihedral.dat",
    "ALLA_length_one.dat", "ALLA_angle_one.dat", "ALLA_dihedral_one.dat",
    "water.gro", "water.xtc", "water.pdb", "water-nobox.pdb", "water.map", "water_double_weight.map", "water-cg.gro", "pbcwater.gro",
    "two.gro", "two.map", "two.itp", "dppc.map",
    "polyethene.gro", "polyethene.bnd", "polyethene.map", "pbcpolyethene.gro",
    "triangle.bnd", "duplicate_atoms.bnd", "twice.cfg", "nosections.cfg",
    "martini3/four.gro", "martini3/four.map", "martini3/four.itp",
    "martini3/naphthalene.gro", "martini3/naphthalene.map", "martini3/naphthalene.bnd",
    "martini3/naphthalene_out.itp",
    "ffout.ff/out.rtp", "ffout.ff/out.r2b",
]


def main():
    for rel in COPIES:
        dst = HERE / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(REF / rel, dst)
    sys.path.insert(0, str(HERE.parent.parent))
    from oracle import bonds, mapping, system
    aa = system.load(HERE / "sugar.gro", HERE / "sugar.xtc")
    cg = mapping.apply(mapping.parse_map(HERE / "sugar.map"), aa)
    terms = bonds.parse_bnd(HERE / "sugar.bnd")
    bonds.measure(terms, cg)
    bonds.invert(terms)
    params = [{"atoms": list(t.atoms), "eqm": t.eqm, "fconst": t.fconst} for t in terms["ALLA"]]
    (HERE / "sugar_params.json").write_text(json.dumps(params, indent=1) + "\n")
    print("wrote", len(COPIES), "copies and sugar_params.json")


if __name__ == "__main__":
    main()
