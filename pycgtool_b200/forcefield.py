"""GROMACS ``.ff`` force-field directory writer (caller of the hot path's results).

Same file formats as ``pycgtool/forcefield.py``: ``<name>.rtp`` (residue building blocks with the
fitted bonded parameters, plus N-/C-/2-terminal variants for residues whose bonds reach into the
neighbouring residue) and ``<name>.r2b``.  The MARTINI parameter files the reference copies next to
them (``martini_v2.2.itp``, ``w.itp``, ``watermodels.dat``) are third-party data and are not shipped
here; pass ``martini_dir`` to have them copied and ``atomtypes.atp`` generated.
"""

import pathlib
import shutil

from .parsers import CFG
from .util import any_starts_with, backup_file, file_write_lines

MARTINI_FILES = ("martini_v2.2.itp", "watermodels.dat", "w.itp")


class ForceField:
    def __init__(self, name, dir_path=pathlib.Path("."), martini_dir=None):
        self.directory = pathlib.Path(dir_path).joinpath(f"ff{name}.ff")
        backup_file(self.directory)
        self.directory.mkdir(parents=True, exist_ok=True)
        with open(self.directory / "forcefield.itp", "w") as itp:
            print(f"#define _FF_PYCGTOOL_{name}", file=itp)
            print('#include "martini_v2.2.itp"', file=itp)
        if martini_dir is not None:
            martini_dir = pathlib.Path(martini_dir)
            for fname in MARTINI_FILES:
                shutil.copyfile(martini_dir / fname, self.directory / fname)
            with CFG(martini_dir / "martini_v2.2.itp") as itp, open(self.directory / "atomtypes.atp", "w") as atp:
                for toks in itp["atomtypes"]:
                    print(" ".join(toks), file=atp)
        with open(self.directory / "forcefield.doc", "w") as doc:
            print(f"PyCGTOOL produced MARTINI force field - {name}", file=doc)

    def write(self, filename, mapping, bonds):
        lines, nterms, cterms = ForceField.write_rtp(mapping, bonds)
        file_write_lines(self.directory / f"{filename}.rtp", lines)
        file_write_lines(self.directory / f"{filename}.r2b", ForceField.write_r2b(nterms, cterms))

    @staticmethod
    def bond_section(bonds, section_header, multiplicity=None):
        lines = ["  [ {0:s} ]".format(section_header)] if bonds else []
        for bond in bonds:
            line = "    " + " ".join("{0:>4s}".format(atom) for atom in bond.atoms)
            line += " {0:12.5f} {1:12.5f}".format(bond.eqm, bond.fconst)
            if multiplicity is not None:
                line += " {0:4d}".format(multiplicity)
            lines.append(line)
        return lines

    @staticmethod
    def needs_terminal_entries(mol_name, bondset):
        bonds = bondset.get_bonds(mol_name, natoms=-1)
        return any_starts_with(bonds, "-"), any_starts_with(bonds, "+")

    @staticmethod
    def write_rtp(mapping, bonds):
        """Lines of the .rtp plus the residues that need N- and C-terminal building blocks."""

        def residue_block(mol_name, mol_mapping, strip=None, prepend=""):
            lines = ["[ {0} ]".format(prepend + mol_name), "  [ atoms ]"]
            for bead in mol_mapping:
                lines.append("    {:>4s} {:>4s} {:3.6f} {:4d}".format(bead.name, bead.type, bead.charge, 0))
            for natoms, (section, multiplicity) in enumerate((("bonds", None), ("angles", None), ("dihedrals", 1)),
                                                             start=2):
                if strip is None:
                    selected = bonds.get_bonds(mol_name, natoms)
                else:
                    selected = bonds.get_bonds(mol_name, natoms,
                                               select=lambda bond: not any_starts_with(bond, strip))
                lines.extend(ForceField.bond_section(selected, section, multiplicity))
            return lines

        n_terms, c_terms = set(), set()
        rtp = ["[ bondedtypes ]", ("{:4d}" * 8).format(1, 1, 1, 1, 1, 1, 0, 0)]
        for mol_name, mol_mapping in mapping.items():
            try:
                rtp.extend(residue_block(mol_name, mol_mapping))
            except KeyError:
                continue
            reaches_back, reaches_forward = ForceField.needs_terminal_entries(mol_name, bonds)
            if reaches_back:
                rtp.extend(residue_block(mol_name, mol_mapping, strip="-", prepend="N"))
                n_terms.add(mol_name)
            if reaches_forward:
                rtp.extend(residue_block(mol_name, mol_mapping, strip="+", prepend="C"))
                c_terms.add(mol_name)
                if reaches_back:
                    rtp.extend(residue_block(mol_name, mol_mapping, strip=("-", "+"), prepend="2"))
        return rtp, n_terms, c_terms

    @staticmethod
    def write_r2b(n_terms, c_terms):
        lines = ["; rtp residue to rtp building block table", ";     main  N-ter C-ter 2-ter"]
        for resname in sorted(n_terms | c_terms):
            nter = ("N" + resname) if resname in n_terms else "-"
            cter = ("C" + resname) if resname in c_terms else "-"
            both = ("2" + resname) if resname in (n_terms & c_terms) else "-"
            lines.append("{0:5s} {0:5s} {1:5s} {2:5s} {3:5s}".format(resname, nter, cter, both))
        return lines
