def alignment_to_fasta(d):
    return "".join(f">{n}\n{s}\n" for n, s in d.items())
