class DnaSequence(str):
    def iter_kmers(self, k, strict=True):
        """all overlapping k-mers (cogent3.core.sequence.Sequence.iter_kmers)"""
        for i in range(len(self) - k + 1):
            yield str.__getitem__(self, slice(i, i + k))
