"""Stand-in for the cogent3 names DBGA imports.  Test infrastructure, see ../README.md."""
from pathlib import Path

IUPAC_DNA = set("ACGTURYMKSWHBVDN?-")


class AlphabetError(Exception):
    pass


class _Seq(str):
    """A sequence that is also its own string (the reference calls str(seq), len(seq), seq[i])."""

    name = None


class SequenceCollection:
    def __init__(self, data, moltype=None):
        if isinstance(data, SequenceCollection):
            data = data.to_dict()
        self._names = list(data.keys())
        self._seqs = []
        for n in self._names:
            s = _Seq(str(data[n]))
            s.name = n
            self._seqs.append(s)
        self.moltype = moltype

    @property
    def names(self):
        return list(self._names)

    @property
    def seqs(self):
        return list(self._seqs)

    @property
    def num_seqs(self):
        return len(self._seqs)

    def iter_seqs(self):
        return iter(self._seqs)

    def to_dict(self):
        return {n: str(s) for n, s in zip(self._names, self._seqs)}

    def to_fasta(self):
        return "".join(f">{n}\n{s}\n" for n, s in zip(self._names, self._seqs))

    def __eq__(self, other):
        if isinstance(other, dict):
            return self.to_dict() == other
        if isinstance(other, SequenceCollection):
            return self.to_dict() == other.to_dict()
        return NotImplemented

    def __repr__(self):
        return f"{type(self).__name__}({self.to_dict()!r})"


class ArrayAlignment(SequenceCollection):
    @property
    def seq_len(self):
        return len(self._seqs[0]) if self._seqs else 0


def _check(data, moltype):
    if moltype == "dna":
        for s in data.values():
            for ch in str(s).upper():
                if ch not in IUPAC_DNA:
                    raise AlphabetError(ch)


def make_unaligned_seqs(data, moltype=None):
    data = dict(data)
    _check(data, moltype)
    return SequenceCollection(data, moltype=moltype)


def make_aligned_seqs(data, moltype=None):
    data = dict(data)
    _check(data, moltype)
    return ArrayAlignment(data, moltype=moltype)


def load_unaligned_seqs(path, moltype=None):
    p = Path(path)
    if not p.is_file():
        raise ValueError(f"{path} is not a file")
    data, name = {}, None
    for line in p.read_text().splitlines():
        line = line.strip()
        if not line:
            continue
        if line.startswith(">"):
            name = line[1:].split()[0]
            data[name] = ""
        else:
            data[name] += line
    return make_unaligned_seqs(data, moltype=moltype)
