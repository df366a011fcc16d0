"""`model.layers` order of the reference's Keras graphs, derived the way Keras derives it.

Why it matters: `keras.Model.save_weights` writes the HDF5 attribute `layer_names` in `model.layers` order and
`load_weights` (load_weights_from_hdf5_group, by_name=False -- what models.py:1217-1218 calls) IGNORES names: it zips
the file's weighted layers with the model's weighted layers positionally.  For a functional model `model.layers` is
not creation order: Network._init_graph_network sorts layers by decreasing depth and, inside a depth, by the order
of a depth-first traversal from the outputs (keras/engine/network.py, Keras 2.2.x [lib]).  In a 5/3 block
(models.py:1248-1270) the two branches are parallel, so the order is a3, c5, b5, d3 = conv2d_1, conv2d_3, conv2d_2,
conv2d_4 while creation order is conv2d_1..4 -- and conv2d_2 / conv2d_3 are both (5,5,128,128): a file written in
creation order would load into real Keras without any shape error and with the two 5x5 kernels of all 18 blocks
swapped.  This module builds the layer graph of DifvdsrDouble.create_model (models.py:1159-1222) with Keras'
auto-names and applies that sort, so the writer emits, and the loader expects, exactly Keras' positions.

(Keras is not installable offline: the algorithm is restated from the Keras 2.2.4 source; parity unpinned.  The
graphs of Difvdsr4 / Difvdsr have no parallel weighted layers, their order is creation order.)"""
from __future__ import annotations


class LayerGraph:
    """Functional-API bookkeeping: layers in creation order with Keras' auto-generated names and inbound layers."""

    def __init__(self):
        self.layers = []       # (name, kind, [inbound layer ids], has_weights)
        self._uid = {}

    def add(self, kind, inbound, name=None, weights=False):
        if name is None:
            self._uid[kind] = self._uid.get(kind, 0) + 1
            name = "%s_%d" % (kind, self._uid[kind])
        self.layers.append((name, kind, list(inbound), weights))
        return len(self.layers) - 1

    def keras_order(self, output):
        """Indices of self.layers in `model.layers` order for Model(inputs, outputs=[layer `output`])."""
        import sys
        layer_indices, finished, post = {}, set(), []

        def build_map(l):
            if l in finished:
                return
            if l not in layer_indices:
                layer_indices[l] = len(layer_indices)          # pre-order position of the traversal from the output
            for inb in self.layers[l][2]:
                build_map(inb)
            finished.add(l)
            post.append(l)

        old = sys.getrecursionlimit()
        sys.setrecursionlimit(max(old, 10 * len(self.layers) + 100))
        try:
            build_map(output)
        finally:
            sys.setrecursionlimit(old)
        depth = {}
        for l in reversed(post):                                # outputs first
            d = depth.setdefault(l, 0)
            for inb in self.layers[l][2]:
                depth[inb] = max(d + 1, depth.get(inb, 0))
        return sorted(post, key=lambda l: (-depth[l], layer_indices[l]))


def difvdsr_double_graph():
    """(LayerGraph, output layer id) of DifvdsrDouble.create_model (models.py:1159-1222); block bodies :1231-1270."""
    g = LayerGraph()
    x = g.add("input", [])
    x = g.add("conv2d", [x], name="level1", weights=True)                 # :1177 (named; does not consume a conv2d uid)

    def block53(x):                                                        # :1248-1270
        ini = g.add("lambda", [x])                                         # resizeBlockLight09
        a = g.add("conv2d", [x], weights=True)                             # k3
        a = g.add("activation", [a])
        a = g.add("conv2d", [a], weights=True)                             # k5
        b = g.add("conv2d", [x], weights=True)                             # k5
        b = g.add("activation", [b])
        b = g.add("conv2d", [b], weights=True)                             # k3
        s = g.add("add", [a, b])
        s = g.add("lambda", [s])                                           # resizeBlockLight01
        return g.add("add", [s, ini])

    def block_light(x):                                                    # :1231-1245
        a = g.add("conv2d", [x], weights=True)
        a = g.add("activation", [a])
        a = g.add("conv2d", [a], weights=True)
        a = g.add("lambda", [a])
        return g.add("add", [a, x])

    for _ in range(16):
        x = block53(x)
    for _ in range(6):
        x = block_light(x)
    x = g.add("lambda", [x])                                               # resizeX4bil :1193
    for _ in range(2):
        x = block53(x)
    out = g.add("conv2d", [x], weights=True)                               # :1199
    # Keras names the InputLayer 'input_1'
    g.layers[0] = ("input_1",) + g.layers[0][1:]
    return g, out


_CACHE = {}


def difvdsr_double_layers():
    """[(name, has_weights)] in Keras `model.layers` order."""
    if "dd" not in _CACHE:
        g, out = difvdsr_double_graph()
        _CACHE["dd"] = [(g.layers[i][0], g.layers[i][3]) for i in g.keras_order(out)]
    return list(_CACHE["dd"])


def difvdsr_double_weighted_order():
    """Names of the weighted layers in the order Keras' load_weights zips them: per 5/3 block a3, c5, b5, d3."""
    return [n for n, w in difvdsr_double_layers() if w]
