"""Point-cloud checkpoints in OpenGaussian's PLY layout: the same vertex properties,
in the same order and with the same conventions as ``GaussianModel.save_ply / load_ply``
(scene/gaussian_model.py:249-298, :305-351), without the ``plyfile`` dependency:

    x y z  nx ny nz  ins_feat_r ins_feat_g ins_feat_b ins_feat_r2 ins_feat_g2 ins_feat_b2
    f_dc_0..2  f_rest_0..44  opacity  scale_0..2  rot_0..3   (all float32)   red green blue (uint8)

``f_dc`` / ``f_rest`` are stored channel-major (``transpose(1, 2).flatten``), the preview colour is
``(ins_feat[:, :3] + 1) / 2 * 255`` clipped, grey (128) where ``sigmoid(opacity) < 0.1``.  Files are
``binary_little_endian 1.0`` with one ``vertex`` element, which is what ``plyfile`` writes by default; the reader
accepts any property order and the scalar types plyfile emits.

The reference fills the vertex table with ``list(map(tuple, attributes))`` (one Python tuple per Gaussian);
here the table is a numpy structured array filled column by column.
"""
import os

import numpy as np
import torch

_PLY_TYPES = {"float": "<f4", "float32": "<f4", "double": "<f8", "float64": "<f8", "uchar": "u1", "uint8": "u1",
              "char": "i1", "int8": "i1", "short": "<i2", "int16": "<i2", "ushort": "<u2", "uint16": "<u2",
              "int": "<i4", "int32": "<i4", "uint": "<u4", "uint32": "<u4"}


def attribute_names(n_dc: int = 3, n_rest: int = 45, n_scale: int = 3, n_rot: int = 4):
    """construct_list_of_attributes (scene/gaussian_model.py:249-262)."""
    names = ["x", "y", "z", "nx", "ny", "nz", "ins_feat_r", "ins_feat_g", "ins_feat_b", "ins_feat_r2", "ins_feat_g2",
             "ins_feat_b2"]
    names += [f"f_dc_{i}" for i in range(n_dc)]
    names += [f"f_rest_{i}" for i in range(n_rest)]
    names.append("opacity")
    names += [f"scale_{i}" for i in range(n_scale)]
    names += [f"rot_{i}" for i in range(n_rot)]
    return names


def vertex_table(xyz, features_dc, features_rest, opacity, scaling, rotation, ins_feat) -> np.ndarray:
    """The structured vertex array of save_ply (:264-296) from the PARAMETER tensors
    (`_xyz [P,3], _features_dc [P,1,3], _features_rest [P,15,3], _opacity [P,1], _scaling [P,3], _rotation [P,4],
    _ins_feat [P,6]`)."""
    def npf(t):
        return t.detach().float().cpu().numpy()

    xyz = npf(xyz)
    f_dc = npf(features_dc.detach().transpose(1, 2).flatten(start_dim=1).contiguous())
    f_rest = npf(features_rest.detach().transpose(1, 2).flatten(start_dim=1).contiguous())
    opac, scale, rot, feat = npf(opacity), npf(scaling), npf(rotation), npf(ins_feat)
    names = attribute_names(f_dc.shape[1], f_rest.shape[1], scale.shape[1], rot.shape[1])
    dtype = [(n, "<f4") for n in names] + [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    el = np.empty(xyz.shape[0], dtype=dtype)
    cols = np.concatenate([xyz, np.zeros_like(xyz), feat, f_dc, f_rest, opac.reshape(-1, 1), scale, rot], axis=1)
    for i, n in enumerate(names):
        el[n] = cols[:, i]
    vis = np.clip((feat + 1) / 2 * 255, 0, 255)
    ignored = (1.0 / (1.0 + np.exp(-opac.reshape(-1)))) < 0.1
    for k, n in enumerate(("red", "green", "blue")):
        c = vis[:, k].copy()
        c[ignored] = 128
        el[n] = c.astype(np.uint8)          # float -> uint8 truncation, as the tuple assignment does
    return el


def write_ply(path: str, el: np.ndarray) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    inv = {"<f4": "float", "|u1": "uchar", "u1": "uchar"}
    lines = ["ply", "format binary_little_endian 1.0", f"element vertex {el.shape[0]}"]
    for n in el.dtype.names:
        lines.append(f"property {inv[el.dtype[n].str]} {n}")
    lines.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(lines) + "\n").encode("ascii"))
        f.write(el.tobytes())


def save_ply(path: str, pc, save_q=()) -> None:
    """GaussianModel.save_ply (:264-298) for any object holding the parameter tensors."""
    feat = pc._ins_feat_q if "ins_feat" in save_q else pc._ins_feat
    write_ply(path, vertex_table(pc._xyz, pc._features_dc, pc._features_rest, pc._opacity, pc._scaling, pc._rotation, feat))


def read_ply(path: str) -> np.ndarray:
    """Vertex element of a binary little-endian (or ascii) PLY as a numpy structured array."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, count, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    count = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError("list properties are not supported in the vertex element")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "binary_little_endian":
            return np.frombuffer(f.read(count * np.dtype(props).itemsize), dtype=props, count=count)
        if fmt == "ascii":
            rows = np.loadtxt(f, max_rows=count, ndmin=2)
            el = np.empty(count, dtype=props)
            for i, (n, _) in enumerate(props):
                el[n] = rows[:, i]
            return el
        raise ValueError(f"{path}: unsupported PLY format {fmt}")


def load_ply(path: str, max_sh_degree: int = 3, device="cuda") -> dict:
    """GaussianModel.load_ply (:305-351): returns the parameter tensors (float32, on `device`) keyed like the
    attributes of GaussianModel (`_xyz, _features_dc [P,1,3], _features_rest [P,15,3], _opacity [P,1], _scaling,
    _rotation, _ins_feat`)."""
    el = read_ply(path)
    names = el.dtype.names

    def cols(prefix_names):
        return np.stack([np.asarray(el[n], dtype=np.float32) for n in prefix_names], axis=1)

    xyz = cols(["x", "y", "z"])
    ins_feat = cols(["ins_feat_r", "ins_feat_g", "ins_feat_b", "ins_feat_r2", "ins_feat_g2", "ins_feat_b2"])
    opac = cols(["opacity"])
    f_dc = cols(["f_dc_0", "f_dc_1", "f_dc_2"])[:, :, None]                                  # [P, 3, 1]
    rest = sorted((n for n in names if n.startswith("f_rest_")), key=lambda n: int(n.split("_")[-1]))
    assert len(rest) == 3 * (max_sh_degree + 1) ** 2 - 3
    f_rest = cols(rest).reshape(xyz.shape[0], 3, (max_sh_degree + 1) ** 2 - 1)                # [P, 3, 15]
    scales = cols(sorted((n for n in names if n.startswith("scale_")), key=lambda n: int(n.split("_")[-1])))
    rots = cols(sorted((n for n in names if n.startswith("rot")), key=lambda n: int(n.split("_")[-1])))
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float, device=device)  # noqa: E731
    return {"_xyz": t(xyz), "_features_dc": t(f_dc).transpose(1, 2).contiguous(),
            "_features_rest": t(f_rest).transpose(1, 2).contiguous(), "_opacity": t(opac), "_scaling": t(scales),
            "_rotation": t(rots), "_ins_feat": t(ins_feat)}
