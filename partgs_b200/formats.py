"""On-disk formats of PartGS: surfel PLY files and training checkpoints.

Host-side code (the reference's is Python + the third-party ``plyfile`` package, absent from this image; its
pinned version is not recorded in the reference tree — ``environment.yml`` only names it).  What is restated here is
the published PLY 1.0 format as ``plyfile`` emits it for the reference's call sites:

* ``GaussianModel._save_ply / _load_ply``            scene/gaussian_model.py:285-372
* ``TwoGaussianModel._save_ply / _load_ply`` (+ ``semantic_i``)  games/block_mesh_splatting/scene/two_gaussian_model.py:257-339
* ``fetchPly / storePly``                            scene/dataset_readers.py:136-159
* ``capture / restore`` tuples                       scene/gaussian_model.py:84-128, two_gaussian_model.py:187-235,
                                                     block_gaussian_model.py:49-96; saved by train.py:142 as
                                                     ``torch.save((tuple, iteration), "chkpnt<iter>.pth")``

The writer produces what ``PlyData([PlyElement.describe(arr, 'vertex')]).write(path)`` produces (header lines, type
names, native little-endian packed records), but fills the record array with one vectorised copy instead of the
reference's per-row ``list(map(tuple, attributes))`` loop (minutes at 1 M surfels).  The reader accepts the three PLY
encodings (ascii, binary little / big endian) and every scalar property type; list properties are parsed for
non-vertex elements (faces) so that meshes exported next to the point cloud can be read back.

Parity note: with ``plyfile`` absent the byte layout is pinned by the PLY specification and by round trips, not by
files written by the reference ("parity unpinned" for this module).
"""
from __future__ import annotations

import os
from collections import OrderedDict, namedtuple
from typing import Dict, Optional, Sequence

import numpy as np

# numpy kind/size -> PLY type name, as plyfile spells them on output
_NP_TO_PLY = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "f4": "float",
              "f8": "double"}
# every spelling the PLY specification (and plyfile) accepts on input
_PLY_TO_NP = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}

BasicPointCloud = namedtuple("BasicPointCloud", ["points", "colors", "normals"])  # utils/graphics_utils.py:17-20


class PlyFormatError(RuntimeError):
    pass


# ---------------------------------------------------------------------------------------------------------------
# generic PLY
# ---------------------------------------------------------------------------------------------------------------
def _ply_type(dt: np.dtype) -> str:
    key = dt.kind + str(dt.itemsize)
    if key not in _NP_TO_PLY:
        raise PlyFormatError(f"dtype {dt} has no PLY scalar type")
    return _NP_TO_PLY[key]


def write_ply(path: str, elements: "OrderedDict[str, np.ndarray] | Dict[str, np.ndarray]", text: bool = False,
              comments: Sequence[str] = ()) -> None:
    """Write structured arrays as PLY elements (scalar properties only), binary little-endian unless ``text``."""
    header = ["ply", "format ascii 1.0" if text else "format binary_little_endian 1.0"]
    header += [f"comment {c}" for c in comments]
    for name, arr in elements.items():
        if arr.dtype.names is None or arr.ndim != 1:
            raise PlyFormatError(f"element {name!r}: expected a 1-D structured array")
        header.append(f"element {name} {arr.shape[0]}")
        for field in arr.dtype.names:
            header.append(f"property {_ply_type(arr.dtype[field])} {field}")
    header.append("end_header")
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)  # the reference's mkdir_p(os.path.dirname(path))
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        for arr in elements.values():
            if text:
                for row in arr:
                    f.write((" ".join(repr(v.item()) if v.dtype.kind == "f" else str(v.item()) for v in row) + "\n")
                            .encode("ascii"))
            else:
                packed = np.dtype([(n, arr.dtype[n].newbyteorder("<")) for n in arr.dtype.names])
                f.write(np.ascontiguousarray(arr.astype(packed, copy=False)).tobytes())


def _parse_header(f):
    if f.readline().strip() != b"ply":
        raise PlyFormatError("not a PLY file (missing magic)")
    fmt = None
    elements = []  # [name, count, [(prop name, scalar np code | (count code, item code))]]
    while True:
        line = f.readline()
        if not line:
            raise PlyFormatError("unexpected end of file inside the PLY header")
        tok = line.decode("ascii", "replace").split()
        if not tok or tok[0] in ("comment", "obj_info"):
            continue
        if tok[0] == "format":
            if len(tok) != 3 or tok[1] not in ("ascii", "binary_little_endian", "binary_big_endian"):
                raise PlyFormatError(f"unsupported PLY format line: {line!r}")
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append([tok[1], int(tok[2]), []])
        elif tok[0] == "property":
            if not elements:
                raise PlyFormatError("property before any element")
            if tok[1] == "list":
                if tok[2] not in _PLY_TO_NP or tok[3] not in _PLY_TO_NP:
                    raise PlyFormatError(f"unknown PLY type in {line!r}")
                elements[-1][2].append((tok[4], (_PLY_TO_NP[tok[2]], _PLY_TO_NP[tok[3]])))
            else:
                if tok[1] not in _PLY_TO_NP:
                    raise PlyFormatError(f"unknown PLY type in {line!r}")
                elements[-1][2].append((tok[2], _PLY_TO_NP[tok[1]]))
        elif tok[0] == "end_header":
            break
        else:
            raise PlyFormatError(f"unknown PLY header keyword {tok[0]!r}")
    if fmt is None:
        raise PlyFormatError("PLY header without a format line")
    return fmt, elements


def _read_list_element(f, fmt, count, props):
    """Elements with list properties (faces): a dict of object arrays / scalar arrays, read row by row."""
    out = {name: [] for name, _ in props}
    bo = "<" if fmt == "binary_little_endian" else ">"
    for _ in range(count):
        if fmt == "ascii":
            tok = f.readline().split()
            pos = 0
            for name, code in props:
                if isinstance(code, tuple):
                    n = int(tok[pos])
                    out[name].append(np.array(tok[pos + 1:pos + 1 + n], dtype=np.float64).astype(code[1]))
                    pos += 1 + n
                else:
                    out[name].append(np.float64(tok[pos]).astype(code))
                    pos += 1
        else:
            for name, code in props:
                if isinstance(code, tuple):
                    n = int(np.frombuffer(f.read(np.dtype(code[0]).itemsize), dtype=bo + code[0])[0])
                    item = np.dtype(bo + code[1])
                    out[name].append(np.frombuffer(f.read(n * item.itemsize), dtype=item).astype(code[1]))
                else:
                    item = np.dtype(bo + code)
                    out[name].append(np.frombuffer(f.read(item.itemsize), dtype=item)[0].astype(code))
    res = {}
    for name, code in props:
        if isinstance(code, tuple):
            col = np.empty(count, dtype=object)
            for i, v in enumerate(out[name]):
                col[i] = v
            res[name] = col
        else:
            res[name] = np.array(out[name], dtype=code)
    return res


def read_ply(path: str) -> "OrderedDict[str, np.ndarray | dict]":
    """All elements of a PLY file, in file order.  Elements with scalar properties only come back as structured
    arrays (native byte order); elements with list properties as ``{property: array}`` dicts."""
    result = OrderedDict()
    with open(path, "rb") as f:
        fmt, elements = _parse_header(f)
        for name, count, props in elements:
            if any(isinstance(code, tuple) for _, code in props):
                result[name] = _read_list_element(f, fmt, count, props)
                continue
            native = np.dtype([(n, code) for n, code in props])
            if fmt == "ascii":
                arr = np.empty(count, dtype=native)
                for i in range(count):
                    tok = f.readline().split()
                    if len(tok) < len(props):
                        raise PlyFormatError(f"element {name!r} row {i}: expected {len(props)} values")
                    for (n, code), t in zip(props, tok):
                        arr[n][i] = float(t) if code[0] == "f" else int(float(t))
            else:
                bo = "<" if fmt == "binary_little_endian" else ">"
                disk = np.dtype([(n, bo + code) for n, code in props])
                raw = f.read(count * disk.itemsize)
                if len(raw) != count * disk.itemsize:
                    raise PlyFormatError(f"element {name!r}: file truncated ({len(raw)} of {count * disk.itemsize} bytes)")
                arr = np.frombuffer(raw, dtype=disk).astype(native)
            result[name] = arr
    return result


# ---------------------------------------------------------------------------------------------------------------
# surfel point clouds
# ---------------------------------------------------------------------------------------------------------------
def _np(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def surfel_attribute_names(n_dc: int, n_rest: int, n_scale: int, n_rot: int, n_semantic: int = 0):
    """``construct_list_of_attributes`` (scene/gaussian_model.py:271-283; two_gaussian_model.py:257-271: the
    ``semantic_i`` columns sit between ``opacity`` and ``scale_i``)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(n_dc)]
    names += [f"f_rest_{i}" for i in range(n_rest)]
    names.append("opacity")
    names += [f"semantic_{i}" for i in range(n_semantic)]
    names += [f"scale_{i}" for i in range(n_scale)]
    names += [f"rot_{i}" for i in range(n_rot)]
    return names


def save_surfel_ply(path: str, xyz, features_dc, features_rest, opacity, scaling, rotation, semantic=None) -> None:
    """``_save_ply`` of both model classes.  Arguments are the *stored* (pre-activation) tensors with the model's
    shapes: ``xyz [P,3]``, ``features_dc [P,1,3]``, ``features_rest [P,(D+1)^2-1,3]``, ``opacity [P,1]``,
    ``scaling [P,2]``, ``rotation [P,4]``, ``semantic [P,S]`` (part models only).  Feature columns are written
    channel-major (``transpose(1,2).flatten(1)``) like the reference.

    The base class writes ``inverse_sigmoid(sigmoid(_opacity))`` (gaussian_model.py:293), the part model the raw
    ``_opacity`` (two_gaussian_model.py:279); callers pass whichever they mirror."""
    xyz = _np(xyz).astype(np.float32).reshape(-1, 3)
    P = xyz.shape[0]
    f_dc = _np(features_dc).astype(np.float32).reshape(P, -1, 3).transpose(0, 2, 1).reshape(P, -1)
    f_rest = _np(features_rest).astype(np.float32).reshape(P, -1, 3).transpose(0, 2, 1).reshape(P, -1)
    cols = [xyz, np.zeros_like(xyz), f_dc, f_rest, _np(opacity).astype(np.float32).reshape(P, 1)]
    n_sem = 0
    if semantic is not None:
        sem = _np(semantic).astype(np.float32).reshape(P, -1)
        n_sem = sem.shape[1]
        cols.append(sem)
    scale = _np(scaling).astype(np.float32).reshape(P, -1)
    rot = _np(rotation).astype(np.float32).reshape(P, -1)
    cols += [scale, rot]
    names = surfel_attribute_names(f_dc.shape[1], f_rest.shape[1], scale.shape[1], rot.shape[1], n_sem)
    table = np.ascontiguousarray(np.concatenate(cols, axis=1))  # [P, n_attr] float32 == the packed records
    assert table.shape[1] == len(names)
    elements = table.view(np.dtype([(n, "<f4") for n in names])).reshape(P)
    write_ply(path, OrderedDict(vertex=elements))


def _sorted_columns(vertex: np.ndarray, prefix: str):
    names = [n for n in vertex.dtype.names if n.startswith(prefix)]
    return sorted(names, key=lambda x: int(x.split("_")[-1]))


def load_surfel_ply(path: str, max_sh_degree: int) -> Dict[str, np.ndarray]:
    """``_load_ply``: float32 arrays with the model's parameter shapes — ``xyz [P,3]``, ``features_dc [P,1,3]``,
    ``features_rest [P,(D+1)^2-1,3]``, ``opacity [P,1]``, ``scaling [P,S]``, ``rotation [P,4]``, ``semantic [P,S]``
    (``[P,0]`` when the file has no ``semantic_i`` columns) and ``active_sh_degree = max_sh_degree``
    (gaussian_model.py:372)."""
    data = read_ply(path)
    vertex = next(iter(data.values()))  # the reference reads plydata.elements[0]
    if isinstance(vertex, dict):
        raise PlyFormatError("the first PLY element has list properties; expected the surfel table")
    for need in ("x", "y", "z", "opacity", "f_dc_0", "f_dc_1", "f_dc_2"):
        if need not in vertex.dtype.names:
            raise PlyFormatError(f"surfel PLY lacks property {need!r}")
    P = vertex.shape[0]
    f32 = np.float32
    xyz = np.stack([vertex["x"], vertex["y"], vertex["z"]], axis=1).astype(f32)
    opacity = vertex["opacity"].astype(f32)[:, None]
    features_dc = np.stack([vertex["f_dc_0"], vertex["f_dc_1"], vertex["f_dc_2"]], axis=1).astype(f32)[:, None, :]
    rest_names = _sorted_columns(vertex, "f_rest_")
    n_rest = (max_sh_degree + 1) ** 2 - 1
    if len(rest_names) != 3 * n_rest:  # the reference asserts (gaussian_model.py:340)
        raise PlyFormatError(f"{len(rest_names)} f_rest columns, expected {3 * n_rest} for SH degree {max_sh_degree}")
    rest = np.stack([vertex[n] for n in rest_names], axis=1).astype(f32) if rest_names else np.zeros((P, 0), f32)
    features_rest = np.ascontiguousarray(rest.reshape(P, 3, n_rest).transpose(0, 2, 1))
    scale_names = _sorted_columns(vertex, "scale_")
    rot_names = _sorted_columns(vertex, "rot")
    sem_names = _sorted_columns(vertex, "semantic_")

    def table(names):
        return np.stack([vertex[n] for n in names], axis=1).astype(f32) if names else np.zeros((P, 0), f32)

    return {"xyz": xyz, "features_dc": np.ascontiguousarray(features_dc), "features_rest": features_rest,
            "opacity": opacity, "scaling": table(scale_names), "rotation": table(rot_names),
            "semantic": table(sem_names), "active_sh_degree": int(max_sh_degree)}


def fetch_ply(path: str) -> BasicPointCloud:
    """``fetchPly`` (scene/dataset_readers.py:136-142): positions, colours / 255, normals of an input point cloud."""
    v = read_ply(path)["vertex"]
    positions = np.vstack([v["x"], v["y"], v["z"]]).T
    colors = np.vstack([v["red"], v["green"], v["blue"]]).T / 255.0
    normals = np.vstack([v["nx"], v["ny"], v["nz"]]).T
    return BasicPointCloud(points=positions, colors=colors, normals=normals)


def store_ply(path: str, xyz, rgb) -> None:
    """``storePly`` (scene/dataset_readers.py:144-159): float positions, zero normals, uchar colours."""
    xyz = _np(xyz)
    rgb = _np(rgb)
    dtype = [("x", "f4"), ("y", "f4"), ("z", "f4"), ("nx", "f4"), ("ny", "f4"), ("nz", "f4"), ("red", "u1"),
             ("green", "u1"), ("blue", "u1")]
    el = np.empty(xyz.shape[0], dtype=dtype)
    for i, n in enumerate("xyz"):
        el[n] = xyz[:, i]
        el["n" + n] = 0
    for i, n in enumerate(("red", "green", "blue")):
        el[n] = rgb[:, i]  # same C cast as numpy's tuple assignment in the reference
    write_ply(path, OrderedDict(vertex=el))


# ---------------------------------------------------------------------------------------------------------------
# checkpoints
# ---------------------------------------------------------------------------------------------------------------
# Field order of the tuples the three model classes hand to torch.save.
CAPTURE_FIELDS = {
    # scene/gaussian_model.py:84-98
    "gaussian": ("active_sh_degree", "_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity",
                 "max_radii2D", "xyz_gradient_accum", "denom", "optimizer_state", "spatial_lr_scale"),
    # games/block_mesh_splatting/scene/two_gaussian_model.py:187-202
    "part": ("active_sh_degree", "_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity",
             "_semantic", "max_radii2D", "xyz_gradient_accum", "denom", "optimizer_state", "spatial_lr_scale"),
    # games/block_mesh_splatting/scene/block_gaussian_model.py:49-69
    "block": ("active_sh_degree", "_features_dc", "_features_rest", "_alpha", "_scale", "sq_r", "sq_s", "sq_t",
              "sq_occ", "sq_eps", "faces", "sq_eta", "sq_omega", "max_radii2D", "xyz_gradient_accum", "denom",
              "optimizer_state", "spatial_lr_scale"),
}
# TwoGaussianModel.restore unpacks positions 7 and 8 as (max_radii2D, _semantic) although capture wrote
# (_semantic, max_radii2D) (two_gaussian_model.py:204-231) — a reference bug that swaps the two on resume.
# `unpack_checkpoint(kind="part")` follows capture's order, i.e. what was actually stored.


def detect_checkpoint_kind(model_args) -> str:
    n = len(model_args)
    for kind, fields in CAPTURE_FIELDS.items():
        if len(fields) == n:
            return kind
    raise ValueError(f"checkpoint tuple of length {n} matches none of {sorted((k, len(v)) for k, v in CAPTURE_FIELDS.items())}")


def pack_checkpoint(state: dict, kind: str) -> tuple:
    """dict (keys of ``CAPTURE_FIELDS[kind]``) -> the tuple ``capture()`` / ``capture_block()`` returns."""
    fields = CAPTURE_FIELDS[kind]
    missing = [f for f in fields if f not in state]
    if missing:
        raise KeyError(f"checkpoint state lacks {missing}")
    return tuple(state[f] for f in fields)


def unpack_checkpoint(model_args, kind: Optional[str] = None) -> dict:
    kind = kind or detect_checkpoint_kind(model_args)
    fields = CAPTURE_FIELDS[kind]
    if len(model_args) != len(fields):
        raise ValueError(f"{kind!r} checkpoints hold {len(fields)} fields, got {len(model_args)}")
    out = dict(zip(fields, model_args))
    out["kind"] = kind
    return out


def checkpoint_path(model_path: str, iteration: int) -> str:
    return os.path.join(model_path, "chkpnt" + str(iteration) + ".pth")  # train.py:61,142


def save_checkpoint(model_path: str, iteration: int, state: dict, kind: str) -> str:
    """``torch.save((gaussians.capture(), iteration), model_path + "/chkpnt<iter>.pth")`` (train.py:142)."""
    import torch

    path = checkpoint_path(model_path, iteration)
    os.makedirs(model_path, exist_ok=True)
    torch.save((pack_checkpoint(state, kind), iteration), path)
    return path


def load_checkpoint(path: str, kind: Optional[str] = None, map_location=None):
    """-> (state dict, iteration); ``(model_params, first_iter) = torch.load(...)`` of train.py:61."""
    import torch

    model_args, iteration = torch.load(path, map_location=map_location, weights_only=False)
    return unpack_checkpoint(model_args, kind), iteration
