"""Host-side helpers shared by the mapping / bond plans and the text writers.

Counterparts of the helpers in ``pycgtool/util.py`` that the hot path's callers
use: the bond-graph walk that generates angles and dihedrals, the sliding
residue window, circular statistics for ad-hoc arrays, GROMACS-style file
backup and the tolerant text comparison used by the tests.
"""

import itertools
import logging
import os
import pathlib
import random

import numpy as np

logger = logging.getLogger(__name__)


class EmptyIterableError(ValueError):
    pass


def circular_mean(values):
    """Mean direction of angles in radians, NaN-skipping (``util.py:24-37``).

    Host helper for ad-hoc arrays; trajectories measured through ``BondSet.apply``
    get their circular statistics from the GPU accumulators instead.
    """
    values = np.asarray(values)
    return np.arctan2(np.nanmean(np.sin(values)), np.nanmean(np.cos(values)))


def circular_variance(values):
    """Mean squared angular distance from the circular mean (``util.py:40-53``)."""
    values = np.asarray(values)
    delta = values - circular_mean(values)
    delta = delta - 2 * np.pi * np.rint(delta / (2 * np.pi))
    return np.nanmean(delta * delta)


def extend_graph_chain(extend, pairs):
    """Lengthen every chain in ``extend`` by one bonded neighbour.

    ``pairs`` are the edges of an undirected graph; each chain is tried at its
    tail and then, reversed, at its head.  Results keep the order and orientation
    in which they are first found; a chain equal to an earlier one read backwards
    is dropped.  A tail node written ``+X`` also continues through X's edges into
    the next residue (GROMACS .rtp convention).  Same contract as ``util.py:56-104``;
    the order matters because generated angles are appended to the bond list in it.
    """
    grown = []
    seen = set()

    def offer(candidate):
        if candidate not in seen and candidate[::-1] not in seen:
            seen.add(candidate)
            grown.append(candidate)

    def neighbours(node, prefix=""):
        for left, right in pairs:
            if node == left:
                yield prefix + right if prefix else right
            elif node == right:
                yield prefix + left if prefix else left

    for chain in extend:
        chain = tuple(chain)
        for oriented in (chain, chain[::-1]):
            tail = oriented[-1]
            for other in neighbours(tail):
                if other not in oriented:
                    offer(oriented + (other,))
            if isinstance(tail, str) and tail.startswith("+"):
                for other in neighbours(tail.strip("+"), prefix="+"):
                    if other not in oriented:
                        offer(oriented + (other,))
    return grown


def sliding(vals):
    """Yield ``(previous, current, next)`` over an iterable, ``None`` at the ends."""
    iterator = iter(vals)
    try:
        current = next(iterator)
    except StopIteration:
        raise EmptyIterableError("Cannot make sliding window over empty iterable")
    previous = None
    for upcoming in iterator:
        yield previous, current, upcoming
        previous, current = current, upcoming
    yield previous, current, None


def transpose_and_sample(sequence, n=None):
    """Rows of the transposed sequence with only finite entries, down-sampled to ``n``."""
    rows = [row for row in zip(*sequence) if np.isfinite(row).all()]
    if n is not None and len(rows) > n:
        rows = random.sample(rows, n)
    return rows


def dir_up(name, n=1):
    for _ in range(n):
        name = os.path.dirname(name)
    return name


def backup_file(path):
    """Move an existing file aside as ``#name.N#`` (GROMACS convention); returns the new name."""
    path = pathlib.Path(path)
    target, counter = path, 1
    while target.exists():
        target = path.parent / f"#{path.name}.{counter}#"
        counter += 1
    if target != path:
        path.rename(target)
        logger.warning("Existing file %s backed up as %s", path.name, target.name)
    return target


def file_write_lines(filename, lines=None, backup=True, append=False):
    if backup and not append:
        backup_file(filename)
    with open(filename, "a" if append else "w") as handle:
        for line in lines or ():
            print(line, file=handle)


def number_or_string(string):
    try:
        number = float(string)
    except ValueError:
        return string
    return int(number) if number == int(number) else number


def cmp_whitespace_float(ref_lines, test_lines, rtol=0.01, verbose=False):
    """Compare line iterables token by token; floats may differ by ``rtol`` relative."""
    failures = []
    for lineno, (ref, test) in enumerate(itertools.zip_longest(ref_lines, test_lines)):
        if ref is None or test is None:
            failures.append((lineno, ref, test))
            continue
        if ref == test:
            continue
        pairs = itertools.zip_longest(map(number_or_string, ref.split()), map(number_or_string, test.split()))
        for a, b in pairs:
            if a == b:
                continue
            try:
                bad = abs(a - b) > abs(a) * rtol
            except TypeError:
                bad = True
            if bad:
                failures.append((lineno, ref, test))
    if verbose and failures:
        print("Lines fail comparison:")
        for lineno, ref, test in failures:
            print(f"Line {lineno}\nRef:  {ref}\nTest: {test}")
    return not failures


def cmp_file_whitespace_float(ref_filename, test_filename, rtol=0.01, verbose=False):
    with open(ref_filename) as ref, open(test_filename) as test:
        return cmp_whitespace_float(ref.readlines(), test.readlines(), rtol=rtol, verbose=verbose)


def any_starts_with(iterable, char):
    if isinstance(iterable, str):
        return iterable.startswith(char)
    return any(any_starts_with(item, char) for item in iterable)


def compare_trajectories(*trajectory_files, topology_file=None, rtol=0.001):
    """Compare trajectories for equality (``pycgtool/util.py:268-318``): same atom count, coordinates within
    ``rtol``, same box.  Returns True, raises ``ValueError`` describing the first difference."""
    from . import formats
    top = None if topology_file is None else formats.load(topology_file).topology
    ref = formats.load(trajectory_files[0], top=top)
    for other_file in trajectory_files[1:]:
        try:
            other = formats.load(other_file, top=top)
        except ValueError as exc:
            raise ValueError("Topology does not match") from exc
        if other.n_atoms != ref.n_atoms or other.n_frames != ref.n_frames:
            raise ValueError("Number of atoms or frames does not match")
        if not np.allclose(other.xyz, ref.xyz, rtol=rtol):
            raise ValueError("Coordinates do not match")
        a, b = ref.unitcell_lengths, other.unitcell_lengths
        if (a is None) != (b is None) or (a is not None and not np.allclose(a, b, rtol=rtol)):
            raise ValueError("Unitcell does not match")
    return True
