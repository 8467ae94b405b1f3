class DistanceMatrix(dict):
    pass
