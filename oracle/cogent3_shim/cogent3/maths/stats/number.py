from collections import Counter


class CategoryCounter(Counter):
    pass
