def tree_align(*a, **k):
    raise NotImplementedError("MSA is out of scope (SURVEY.md section 2)")
