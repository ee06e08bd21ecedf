def nodes_to_centrality(*a, **k):
    raise NotImplementedError
