class ClusterEnvironment:
    pass
