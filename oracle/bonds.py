"""Oracle for bonded-term definition, measurement and inversion.  TEST INFRASTRUCTURE ONLY.

Restates ``pycgtool/bondset.py:149-234`` (``.bnd`` parse + automatic angle /
dihedral generation), ``pycgtool/util.py:56-104`` (``extend_graph_chain``),
``pycgtool/bondset.py:496-531`` (``BondSet.apply``: name -> index resolution with
``+``/``-`` neighbour prefixes, then the mdtraj geometry call) and
``bondset.py:533-539`` (``boltzmann_invert`` over all bonds).
"""

import dataclasses
import typing

import numpy as np

from . import fileio, geometry, stats


@dataclasses.dataclass
class Term:
    atoms: typing.Tuple[str, ...]
    form: str
    values: typing.Any = dataclasses.field(default_factory=list)
    indices: typing.Any = None     # (n_inst, natoms) int, set by measure()
    eqm: typing.Optional[float] = None
    fconst: typing.Optional[float] = None


def grow_chains(chains, edges):
    """``util.extend_graph_chain`` (util.py:56-104): extend every chain by one
    bonded neighbour at either end, keeping first-seen orientation and order and
    dropping reversed duplicates.  A trailing ``+name`` node continues into the
    next residue (util.py:82-96)."""
    found = []

    def keep(item):
        if item not in found and item[::-1] not in found:
            found.append(item)

    for chain in chains:
        for _ in range(2):
            tail = chain[-1]
            for p, q in edges:
                if tail == p and q not in chain:
                    keep(chain + (q,))
                elif tail == q and p not in chain:
                    keep(chain + (p,))
            if isinstance(tail, str) and tail.startswith("+"):
                bare = tail.strip("+")
                for p, q in edges:
                    if bare == p and "+" + q not in chain:
                        keep(chain + ("+" + q,))
                    elif bare == q and "+" + p not in chain:
                        keep(chain + ("+" + p,))
            chain = chain[::-1]
    return found


def parse_bnd(path, generate_angles=True, generate_dihedrals=False, **form_options):
    """``BondSet.__init__`` (bondset.py:149-222) -> ``{molecule: [Term]}``."""
    forms = stats.default_forms(**form_options)
    out = {}
    for mol, rows in fileio.read_cfg(path).items():
        terms = []
        for names in rows:
            if len(set(names)) != len(names):
                raise ValueError(f"Defined bond '{names}' contains duplicate atoms")
            terms.append(Term(tuple(names), forms[len(names)]))
        if not any(len(t.atoms) > 2 for t in terms):
            pairs = [t.atoms for t in terms if len(t.atoms) == 2]
            triples = grow_chains(pairs, pairs)
            quads = grow_chains(triples, pairs)
            if generate_angles:
                terms.extend(Term(a, forms[3]) for a in triples)
            if generate_dihedrals:
                terms.extend(Term(d, forms[4]) for d in quads)
        out[mol] = terms
    return out


def resolve_indices(term, mol, sysm):
    """Index resolution of ``BondSet.apply`` (bondset.py:509-528): one instance per
    residue named ``mol``; ``-``/``+`` prefixed names resolve in the previous /
    next residue whatever its name; a missing neighbour skips the instance."""
    rows = []
    residues = sysm.residues
    for r, res in enumerate(residues):
        if res.name != mol:
            continue
        row = []
        for name in term.atoms:
            target = res
            if name[0] == "-":
                target = residues[r - 1] if r > 0 else None
            elif name[0] == "+":
                target = residues[r + 1] if r + 1 < len(residues) else None
            if target is None:
                row = None
                break
            row.append(target.index_of(name.lstrip("-+")))
        if row is not None:
            rows.append(row)
    return np.array(rows, dtype=np.int32).reshape(-1, len(term.atoms))


def measure(mols, sysm):
    """``BondSet.apply`` (bondset.py:496-531): ``values`` = frame-major 1-D float32."""
    for mol, terms in mols.items():
        for term in terms:
            term.indices = resolve_indices(term, mol, sysm)
            n = len(term.atoms)
            if len(term.indices) == 0:
                # mdtraj raises on an empty index list; the oracle records "no values"
                term.values = np.zeros(0, dtype=np.float32)
                continue
            term.values = geometry.MEASURE[n](sysm.xyz, term.indices, sysm.box_vectors).flatten()


def invert(mols, temp=310.0):
    """``BondSet.boltzmann_invert`` (bondset.py:533-539)."""
    for terms in mols.values():
        for term in terms:
            try:
                term.eqm, term.fconst = stats.boltzmann_invert(term.values, len(term.atoms), term.form, temp)
            except ValueError:
                pass
