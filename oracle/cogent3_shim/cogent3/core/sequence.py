class DnaSequence(str):
    pass
