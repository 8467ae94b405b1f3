def alignment_to_fasta(d, block_size=60, order=None):
    """shim of cogent3.format.fasta.alignment_to_fasta: FASTA text, rows wrapped at block_size columns"""
    out = []
    for n in (order or list(d)):
        s = str(d[n])
        out.append(f">{n}\n")
        out.extend(s[i:i + block_size] + "\n" for i in range(0, len(s), block_size))
    return "".join(out)
