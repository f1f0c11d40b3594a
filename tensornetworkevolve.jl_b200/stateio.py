"""State I/O: a PEPS on disk is its flat ``variables`` vector (src/peps.jl:202-211: vertex
tensors in vertex order, column-major, then the Vidal bond vectors in ``virtual_labels`` order) plus the label
metadata the constructors derive everything else from (src/peps.jl:69-82, 142-151).  The reference has no I/O of
its own; this is the format a time evolution checkpoints to, and the layout ``load_variables!`` consumes as is.

File: a numpy ``.npz`` (no pickling) with

    format            "tne-peps-state", version 1
    kind              "simple" | "vidal"
    nflavor, Dmax, eps
    vertex_labels     flat int64 + vertex_rank (one entry per vertex; physical label first, src/peps.jl:253)
    virtual_labels    int64
    shapes            flat int64 of every tensor of ``alltensors`` (ranks: vertex_rank, then 1 per bond vector)
    variables         complex128, ``sum(prod(shape))`` elements
    time, step        optional scalars of the evolution that wrote the file

``pack_state`` / ``unpack_state`` / ``write_state`` / ``read_state`` are host-only (numpy); ``save_peps`` / ``load_peps``
move the vector between the file and device tensors (one ``tne_download`` / ``tne_upload`` of the whole vector)."""
import numpy as np

FORMAT = "tne-peps-state"
VERSION = 1


def pack_state(kind, vertex_labels, virtual_labels, tensors, Dmax, eps, nflavor=2, time=None, step=None):
    """Host arrays (Julia shapes, any memory order) -> dict of plain numpy arrays in the file layout."""
    if kind not in ("simple", "vidal"):
        raise ValueError("kind must be 'simple' or 'vidal'")
    nv = len(vertex_labels)
    nexpect = nv + (len(virtual_labels) if kind == "vidal" else 0)
    if len(tensors) != nexpect:
        raise ValueError("expected %d tensors (alltensors order), got %d" % (nexpect, len(tensors)))
    tensors = [np.asarray(t, dtype=np.complex128) for t in tensors]
    for t, vl in zip(tensors[:nv], vertex_labels):
        if t.ndim != len(vl):
            raise ValueError("tensor rank %d does not match its %d labels" % (t.ndim, len(vl)))
    for t in tensors[nv:]:
        if t.ndim != 1:
            raise ValueError("bond tensors are vectors (src/peps.jl:131)")
    d = {
        "format": np.array(FORMAT), "version": np.array(VERSION, dtype=np.int64), "kind": np.array(kind),
        "nflavor": np.array(nflavor, dtype=np.int64), "Dmax": np.array(Dmax, dtype=np.int64), "eps": np.array(eps, dtype=np.float64),
        "vertex_rank": np.array([len(vl) for vl in vertex_labels], dtype=np.int64),
        "vertex_labels": np.array([l for vl in vertex_labels for l in vl], dtype=np.int64),
        "virtual_labels": np.array(list(virtual_labels), dtype=np.int64),
        "shapes": np.array([s for t in tensors for s in t.shape], dtype=np.int64),
        "variables": (np.concatenate([t.reshape(-1, order="F") for t in tensors]) if tensors else np.zeros(0, np.complex128)),
    }
    if time is not None:
        d["time"] = np.array(time, dtype=np.float64)
    if step is not None:
        d["step"] = np.array(step, dtype=np.int64)
    return d


def unpack_state(d):
    """Inverse of ``pack_state``: ``(meta, tensors)`` with Fortran-ordered complex128 tensors.  Validates the layout."""
    if str(d["format"]) != FORMAT:
        raise ValueError("not a %s file" % FORMAT)
    if int(d["version"]) != VERSION:
        raise ValueError("unsupported %s version %d" % (FORMAT, int(d["version"])))
    kind = str(d["kind"])
    rank = [int(r) for r in d["vertex_rank"]]
    flat = [int(l) for l in d["vertex_labels"]]
    if sum(rank) != len(flat):
        raise ValueError("vertex_labels does not match vertex_rank")
    vertex_labels, k = [], 0
    for r in rank:
        vertex_labels.append(flat[k:k + r])
        k += r
    virtual_labels = [int(l) for l in d["virtual_labels"]]
    ranks = rank + ([1] * len(virtual_labels) if kind == "vidal" else [])
    shp = [int(s) for s in d["shapes"]]
    if sum(ranks) != len(shp):
        raise ValueError("shapes does not match the tensor ranks")
    var = np.asarray(d["variables"], dtype=np.complex128)
    tensors, k, o = [], 0, 0
    for r in ranks:
        shape = tuple(shp[k:k + r])
        n = int(np.prod(shape)) if shape else 1
        if o + n > var.size:
            raise ValueError("variables is shorter than the shapes require")
        tensors.append(var[o:o + n].reshape(shape, order="F"))
        k += r
        o += n
    if o != var.size:
        raise ValueError("variables has %d elements, the shapes need %d" % (var.size, o))
    meta = dict(kind=kind, nflavor=int(d["nflavor"]), Dmax=int(d["Dmax"]), eps=float(d["eps"]),
                vertex_labels=vertex_labels, virtual_labels=virtual_labels,
                time=float(d["time"]) if "time" in d else None, step=int(d["step"]) if "step" in d else None)
    return meta, tensors


def write_state(path, packed):
    with open(path, "wb") as f:
        np.savez(f, **packed)


def read_state(path):
    with np.load(path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def save_peps(path, peps, time=None, step=None):
    """Checkpoint a device PEPS: ONE device->host copy of ``variables(peps)`` (Vidal: incl. the bond vectors)."""
    from . import peps as P
    var = P.variables(peps).to_numpy()
    ats = peps.alltensors()
    d = pack_state(peps.kind, peps.vertex_labels, peps.virtual_labels, [np.zeros(t.shape, np.complex128) for t in ats],
                   peps.Dmax, peps.eps, peps.nflavor, time, step)
    d["variables"] = np.asarray(var, dtype=np.complex128).reshape(-1)
    write_state(path, d)
    return path


def load_peps(path, optimizer="greedy", nslices=0, segments=None):
    """Rebuild the PEPS (labels -> contraction codes as in the constructors) and upload its tensors.
    Returns ``(peps, meta)``; the contraction order is not part of the state (src/peps.jl:84-93 derives it)."""
    from . import peps as P
    from .array import B200Array
    meta, tensors = unpack_state(read_state(path))
    dev = [B200Array.from_numpy(np.asfortranarray(t)) for t in tensors]
    nv = len(meta["vertex_labels"])
    p = P.make_peps(meta["kind"], meta["vertex_labels"], dev[:nv], meta["virtual_labels"], meta["Dmax"], meta["eps"],
                    meta["nflavor"], dev[nv:] if meta["kind"] == "vidal" else None, optimizer, nslices, segments)
    return p, meta
