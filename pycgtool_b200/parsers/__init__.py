"""Parsers for the sectioned text formats PyCGTOOL reads (.map, .bnd, .itp)."""

from .cfg import CFG, DuplicateSectionError, NoSectionError
from .itp import ITP

__all__ = ["CFG", "ITP", "DuplicateSectionError", "NoSectionError"]
