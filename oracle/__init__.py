"""CPU oracle for the PyCGTOOL per-frame hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the reference algorithm
(jag1g13/pycgtool 2.0.0 + the geometry kernels of its un-vendored dependency
mdtraj 1.9.7, pinned at ``poetry.lock:391-392``).  It exists so that the CUDA
path in ``pycgtool_b200`` can be checked for parity.  Nothing in the product
package imports it: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may.

Why a restatement and not the reference itself: every reference module on the
path does ``import mdtraj`` (``pycgtool/frame.py:11``, ``mapping.py:13-14``,
``bondset.py:13``, ``util.py:12``) and mdtraj is neither installed nor
installable here (no network), so the reference cannot be imported.

Parity pin (see ``tests/test_oracle_golden.py``): the restatement is checked
against the reference's own golden files --
``tests/data/sugar_out.xtc`` (mapping, bit-exact on all 1001x6x3 floats),
``tests/data/ALLA_{length,angle,dihedral}.dat`` (mdtraj outputs as consumed by
the reference, 5 printed decimals), ``tests/data/ffout.ff/out.rtp`` (current
circular statistics + Boltzmann inversion) and the unit KATs in
``tests/test_mapping.py``, ``tests/test_bondset.py``, ``tests/test_util.py``.
The files are copied to ``tests/golden/`` by ``tests/golden/make_golden.py``.
Unpinned: triclinic minimum-image *wrapping* (no reference fixture triggers an
image shift in a triclinic cell) -- restated from mdtraj's published algorithm.

Layout (each function cites the reference file:line it follows):

* ``fileio``    CFG / GRO / XTC readers (``parsers/cfg.py``; mdtraj formats)
* ``system``    minimal residue/atom container (what ``frame.py`` exposes)
* ``mapping``   ``Mapping.__init__``/``apply`` + ``calc_coords_weight``
* ``geometry``  mdtraj ``compute_distances/angles/dihedrals`` with MIC
* ``bonds``     ``BondSet.__init__/apply/boltzmann_invert`` + graph helpers
* ``stats``     ``util.circular_mean/variance`` + ``functionalforms``
"""
