"""Main branches of a summary table (reference summary_utils.py:7-62, behind summarize(find_main_branch=True)).

Host-side post-processing of the P-row branch table, exactly as in the reference: it is not part of the B200
hot path -- the table has one row per branch and the reference runs this step on the
CPU with networkx too.  The same networkx calls in the same order keep node/edge iteration order, and with it
every tie-break, identical.

Two reference quirks are reproduced on purpose (golden vectors from the unmodified reference pin them,
tests/golden/make_main_branch.py):

* find_main_branch_nx looks the edge weight up under 'branch-distance' (summary_utils.py:7) while
  find_main_branches stores it under networkx's default key 'weight' (summary_utils.py:56), so networkx falls
  back to weight 1 per edge: "longest shortest path" is measured in branches, not in branch_distance.
* find_main_branches marks a row for EVERY edge of the annotated graph (summary_utils.py:59-60 iterate
  h.edges(), not the edges carrying main=True), so `main` is False only for rows shadowed by a later row between
  the same pair of nodes (parallel branches; nx.Graph keeps one edge per pair).
"""
import numpy as np

WEIGHT_KEY = 'branch-distance'


def find_main_branch_nx(g, weight=WEIGHT_KEY, in_place=True):
    """Annotate the edges on the longest shortest path of every connected component of `g` with main=True
    (summary_utils.py:7-29)."""
    import networkx as nx
    if not in_place:
        g = g.copy()
    for nodes in nx.connected_components(g):
        sub = g.subgraph(nodes)
        best, ends = 0, None
        for a, reach in nx.all_pairs_dijkstra_path_length(sub, weight=weight):
            for b, d in reach.items():
                if d is not None and np.isfinite(d) and d >= best:   # the last maximum wins
                    best, ends = d, (a, b)
        walk = nx.shortest_path(sub, source=ends[0], target=ends[1], weight=weight)
        for a, b in zip(walk[:-1], walk[1:]):
            g.edges[a, b]['main'] = True
    return g


def find_main_branches(summary, *, annotate=False):
    """Boolean array with one entry per row of `summary` (columns node_id_src, node_id_dst, branch_distance,
    names before the separator is applied), as summary_utils.py:32-62 returns it.

    The reference's return value depends only on which rows own an edge of the branch graph (see the module
    docstring), so the all-pairs Dijkstra that annotates the graph is skipped unless `annotate` asks for the
    annotated graph as well (then `(is_main, graph)` is returned).
    """
    import networkx as nx
    src = np.asarray(summary['node_id_src']).tolist()
    dst = np.asarray(summary['node_id_dst']).tolist()
    length = np.asarray(summary['branch_distance']).tolist()
    row_of = {(u, v): i for i, (u, v) in enumerate(zip(src, dst))}
    row_of.update({(v, u): i for i, (u, v) in enumerate(zip(src, dst))})
    g = nx.Graph()
    g.add_weighted_edges_from(zip(src, dst, length))
    if annotate:
        find_main_branch_nx(g)
    is_main = np.zeros(len(src), dtype=bool)
    for a, b in g.edges():
        is_main[row_of[(a, b)]] = True
    return (is_main, g) if annotate else is_main
