"""``ITP``: GROMACS topology files grouped by ``[ moleculetype ]``.

Same contract as ``pycgtool/parsers/itp.py:12-62``: ``itp[molecule][section]``
is a list of token tuples for the sections atoms / bonds / angles / dihedrals;
other sections are skipped; a molecule defined twice is an error.
"""

import collections
import logging

from .cfg import CFG, DuplicateSectionError, NoSectionError

logger = logging.getLogger(__name__)

_KEPT = ("atoms", "bonds", "angles", "dihedrals")


class ITP(CFG):
    def _read_file(self, filepath):
        section, molecule = None, None
        with open(filepath) as handle:
            for raw in handle:
                text = self._clean(raw, filepath)
                if not text:
                    continue
                if text[0] == "[":
                    section = text.strip("[ ]")
                    if section in _KEPT:
                        self[molecule].setdefault(section, [])
                    continue
                tokens = tuple(text.split())
                if section == "moleculetype":
                    molecule = tokens[0]
                    if molecule in self:
                        raise DuplicateSectionError(molecule, filepath)
                    self[molecule] = collections.OrderedDict()
                elif section in _KEPT:
                    self[molecule][section].append(tokens)
                elif section is None:
                    raise NoSectionError(filepath)
                else:
                    logger.info("File '%s' contains unexpected section '%s'", filepath, section)
