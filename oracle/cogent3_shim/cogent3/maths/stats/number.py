import math
from collections import Counter


class CategoryCounter(Counter):
    @property
    def entropy(self):
        """Shannon entropy (bits) of the category frequencies (cogent3.maths.stats.number.CategoryCounter)"""
        total = sum(self.values())
        return -sum(c / total * math.log2(c / total) for c in self.values() if c)
