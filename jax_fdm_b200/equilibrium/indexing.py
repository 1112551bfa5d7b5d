"""
Mirror of `jax_fdm.equilibrium.indexing` (indexing.py): element keys -> positions in a structure's canonical
ordering, loop-free (argsort + searchsorted; edge pairs folded to one order-sensitive integer code).
"""
import numpy as np

__all__ = ["indices_from_keys"]


def indices_from_keys(keys_canonical, keys_query):
    """
    The position of each queried key in the canonical ordering: a 1-D array of node / vertex / face keys, or a
    2-D array of edge key pairs (one row per element).  `keys_query`: a single key, an edge pair, or a sequence of
    either.  KeyError when a queried key is absent (the folding base covers the queried keys too, so an absent edge
    can never alias a real one).
    """
    canonical = np.asarray(keys_canonical)
    query = np.asarray(keys_query)
    if canonical.ndim == 2:
        query = query.reshape(-1, canonical.shape[1])
        base = int(canonical.max())
        if query.size:
            base = max(base, int(query.max()))
        base += 1
        canonical = canonical[:, 0].astype(np.int64) * base + canonical[:, 1]
        query = query[:, 0].astype(np.int64) * base + query[:, 1]
    else:
        query = query.reshape(-1)
    order = np.argsort(canonical, kind="stable")
    pos = np.clip(np.searchsorted(canonical[order], query), 0, max(canonical.size - 1, 0))
    found = order[pos] if canonical.size else pos
    if canonical.size == 0 or not np.array_equal(canonical[found], query):
        raise KeyError("a queried key is absent from the canonical ordering")
    return found.astype(np.int64)
