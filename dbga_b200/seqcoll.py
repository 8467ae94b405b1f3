"""Minimal sequence containers standing in for the cogent3 objects the reference passes
around (`SequenceCollection`, the object `make_aligned_seqs` returns).  When cogent3 is
importable the public functions hand out real cogent3 objects instead, so the return
types are the reference's own (debruijn_pairwise.py:543, utils.py:95-101)."""
from __future__ import annotations

from pathlib import Path
from typing import Any, Dict, Iterable, List

IUPAC_DNA = frozenset("ACGTURYMKSWHBVDN?-")

try:  # pragma: no cover - cogent3 is not installed in the build image
    import cogent3 as _c3
    HAVE_COGENT3 = True
except Exception:  # noqa: BLE001
    _c3 = None
    HAVE_COGENT3 = False


class AlphabetError(ValueError):
    """Raised for characters outside the IUPAC DNA alphabet (cogent3 raises its own
    AlphabetError at the same point, tests/test_utils.py:69-72)."""


class Sequence(str):
    name: Any = None

    def __new__(cls, value, name=None):
        obj = super().__new__(cls, value)
        obj.name = name
        return obj


class SequenceCollection:
    """Ordered name -> sequence mapping with the attribute surface DBGA uses:
    .names, .seqs, .num_seqs, .iter_seqs(), .to_dict(), .to_fasta(), == dict."""

    def __init__(self, data: Any, moltype: str | None = None):
        if isinstance(data, SequenceCollection):
            data = data.to_dict()
        self._names: List[Any] = list(data.keys())
        self._seqs: List[Sequence] = [Sequence(str(data[n]), n) for n in self._names]
        self.moltype = moltype

    @property
    def names(self):
        return list(self._names)

    @property
    def seqs(self):
        return list(self._seqs)

    @property
    def num_seqs(self) -> int:
        return len(self._seqs)

    def iter_seqs(self) -> Iterable[Sequence]:
        return iter(self._seqs)

    def to_dict(self) -> Dict[Any, str]:
        return {n: str(s) for n, s in zip(self._names, self._seqs)}

    def to_fasta(self, block_size: int = 60) -> str:
        out = []
        for n, s in zip(self._names, self._seqs):
            out.append(f">{n}\n")
            for i in range(0, len(s), block_size):
                out.append(s[i:i + block_size] + "\n")
        return "".join(out)

    def __eq__(self, other):
        if isinstance(other, dict):
            return self.to_dict() == other
        if isinstance(other, SequenceCollection):
            return self.to_dict() == other.to_dict()
        return NotImplemented

    def __repr__(self):
        return f"{type(self).__name__}({self.to_dict()!r})"


class Alignment(SequenceCollection):
    @property
    def seq_len(self) -> int:
        return len(self._seqs[0]) if self._seqs else 0


def _validate(data: Dict[Any, str], moltype):
    if moltype == "dna":
        for s in data.values():
            bad = set(s) - IUPAC_DNA
            if bad:
                raise AlphabetError(f"characters {sorted(bad)!r} are not in the DNA alphabet")


def make_unaligned_seqs(data, moltype=None):
    if HAVE_COGENT3:  # pragma: no cover
        return _c3.make_unaligned_seqs(data, moltype=moltype)
    data = {k: str(v).upper() for k, v in dict(data).items()}
    _validate(data, moltype)
    return SequenceCollection(data, moltype)


def make_aligned_seqs(data, moltype=None):
    if HAVE_COGENT3:  # pragma: no cover
        return _c3.make_aligned_seqs(data, moltype=moltype)
    data = {k: str(v) for k, v in dict(data).items()}
    _validate(data, moltype)
    return Alignment(data, moltype)


def read_fasta(path) -> Dict[str, str]:
    data: Dict[str, str] = {}
    name = None
    with open(path, "r") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line[0] == ">":
                name = line[1:].split()[0] if len(line) > 1 else ""
                data[name] = ""
            elif name is not None:
                data[name] += line
    return data


def load_unaligned_seqs(path, moltype=None):
    if HAVE_COGENT3:  # pragma: no cover
        return _c3.load_unaligned_seqs(path, moltype=moltype)
    p = Path(path)
    if not p.is_file():
        raise ValueError(f"{path} is not a readable sequence file")
    return make_unaligned_seqs(read_fasta(p), moltype=moltype)


def is_collection(obj) -> bool:
    if isinstance(obj, SequenceCollection):
        return True
    return HAVE_COGENT3 and isinstance(obj, _c3.SequenceCollection)  # pragma: no cover
