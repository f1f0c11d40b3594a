"""Host-side ``Graphs.SimpleGraph`` mirror (Graphs 1.7; used at src/peps.jl:244-267 and
src/tebd.jl:16-18,34-37).  Vertices are 1-based like the reference; ``edges`` iterates
``(src < dst)`` pairs in lexicographic order and ``neighbors`` is ascending -- both orders are
observable: they define the virtual label numbering (src/peps.jl:248-253).
"""


class SimpleGraph:
    def __init__(self, n):
        self.n = int(n)
        self._adj = [set() for _ in range(self.n + 1)]

    def add_edge(self, i, j):
        if i == j or not (1 <= i <= self.n and 1 <= j <= self.n):
            return False
        if j in self._adj[i]:
            return False
        self._adj[i].add(j)
        self._adj[j].add(i)
        return True

    def nv(self):
        return self.n

    def ne(self):
        return sum(len(a) for a in self._adj) // 2

    def edges(self):
        return [(i, j) for i in range(1, self.n + 1) for j in sorted(self._adj[i]) if i < j]

    def neighbors(self, i):
        return sorted(self._adj[i])

    def degree(self, i):
        return len(self._adj[i])

    def vertices(self):
        return list(range(1, self.n + 1))


def graph_from_edges(n, edges):
    g = SimpleGraph(n)
    for i, j in edges:
        g.add_edge(i, j)
    return g


def grid(dims):
    """``Graphs.grid([Lx, Ly])``: vertex ``v = x + (y-1)*Lx`` (1-based, x fastest)."""
    lx, ly = dims
    g = SimpleGraph(lx * ly)
    for y in range(ly):
        for x in range(lx):
            v = x + y * lx + 1
            if x + 1 < lx:
                g.add_edge(v, v + 1)
            if y + 1 < ly:
                g.add_edge(v, v + lx)
    return g


def test_graph():
    """The 5-vertex, 6-edge graph every reference test uses (``test/peps.jl:5-8``)."""
    return graph_from_edges(5, [(1, 2), (1, 3), (2, 4), (2, 5), (3, 4), (3, 5)])
