"""``CFG``: the ``[ section ]`` + token-lines format of .map and .bnd files.

Same contract as the reference parser (``pycgtool/parsers/cfg.py:30-88``):
an ordered mapping ``section -> list of token tuples``; ``;`` starts a comment;
``#include "file"`` merges another file's sections; a repeated section or a
token line before any header is an error.  Usable as a context manager.
"""

import collections
import pathlib


class DuplicateSectionError(KeyError):
    def __init__(self, section, filename):
        super().__init__(f"Section '{section}' appears twice in file '{filename}'.")


class NoSectionError(KeyError):
    def __init__(self, filename):
        super().__init__(f"File '{filename}' contains no '[]' section headers.")


class CFG(collections.OrderedDict):
    def __init__(self, filepath=None):
        super().__init__()
        self.filepath = None
        if filepath is not None:
            self.filepath = pathlib.Path(filepath)
            self._read_file(self.filepath)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_value, traceback):
        return None

    def _clean(self, raw, filepath):
        """Comment-stripped line; an include directive is consumed here and yields ''."""
        text = raw.partition(";")[0].strip()
        if text.startswith("#include"):
            target = text.split()[1].strip('"')
            self.update(type(self)(filepath.parent / target))
            return ""
        return text

    def _read_file(self, filepath):
        section = None
        with open(filepath) as handle:
            for raw in handle:
                text = self._clean(raw, filepath)
                if not text:
                    continue
                if text[0] == "[":
                    section = text.strip("[ ]")
                    if section in self:
                        raise DuplicateSectionError(section, filepath)
                    self[section] = []
                elif section in self:
                    self[section].append(tuple(text.split()))
                else:
                    raise NoSectionError(filepath)
