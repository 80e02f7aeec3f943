"""3DGS-style ``.ply`` interchange of Gaussians: ``export_ply`` writes what the reference's
src/model/ply_export.py:26-63 writes -- same signature, same vertex properties in the same order (x y z, nx ny nz, f_dc_0..2,
opacity as a logit, scale_0..2 as logs, rot_0..3 as a wxyz quaternion), positions and rotations taken into the first context
camera's orientation, DC band only -- as a binary little-endian PLY (plyfile's default), without the ``plyfile`` dependency.
``load_ply`` reads such a file (or any 3DGS ply with those properties) back, and ``gaussians_from_ply`` turns it into the
decoder's ``Gaussians`` (covariances R S S^T R^T), which is how a real scene becomes a fixture for tests and bench runs.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch
from torch import Tensor

from .types import Gaussians


def construct_list_of_attributes(num_rest: int) -> list[str]:
    attributes = ["x", "y", "z", "nx", "ny", "nz"]
    for i in range(3):
        attributes.append(f"f_dc_{i}")
    for i in range(num_rest):
        attributes.append(f"f_rest_{i}")
    attributes.append("opacity")
    for i in range(3):
        attributes.append(f"scale_{i}")
    for i in range(4):
        attributes.append(f"rot_{i}")
    return attributes


def _vertex_table(extrinsics: Tensor, means: Tensor, scales: Tensor, rotations: Tensor, harmonics: Tensor, opacities: Tensor) -> np.ndarray:
    """[gaussian, 17] float32 rows in property order (ply_export.py:36-60)."""
    from scipy.spatial.transform import Rotation as R
    view_rotation = extrinsics[:3, :3].inverse()
    means = torch.einsum("ij,...j->...i", view_rotation, means)
    rot = R.from_quat(rotations.detach().cpu().numpy()).as_matrix()
    rot = view_rotation.detach().cpu().numpy() @ rot
    x, y, z, w = R.from_matrix(rot).as_quat().T
    rot = np.stack((w, x, y, z), axis=-1)
    cols = (
        means.detach().cpu().numpy(),
        torch.zeros_like(means).detach().cpu().numpy(),
        harmonics[..., 0].detach().cpu().contiguous().numpy(),
        torch.logit(opacities[..., None]).detach().cpu().numpy(),
        scales.log().detach().cpu().numpy(),
        rot,
    )
    return np.concatenate(cols, axis=1).astype(np.float32)


def export_ply(extrinsics: Tensor, means: Tensor, scales: Tensor, rotations: Tensor, harmonics: Tensor, opacities: Tensor, path: Path) -> None:
    """extrinsics [4,4]; means [g,3]; scales [g,3]; rotations [g,4] xyzw; harmonics [g,3,d_sh]; opacities [g]."""
    table = _vertex_table(extrinsics, means, scales, rotations, harmonics, opacities)
    names = construct_list_of_attributes(0)
    assert table.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {table.shape[0]}\n" + \
             "".join(f"property float {n}\n" for n in names) + "end_header\n"
    path = Path(path)
    path.parent.mkdir(exist_ok=True, parents=True)
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(np.ascontiguousarray(table, dtype="<f4").tobytes())


def load_ply(path: Path) -> dict:
    """-> dict(means [g,3], scales [g,3], rotations [g,4] xyzw, harmonics [g,3,d_sh], opacities [g]) as float32 tensors.
    Reads binary little-endian and ascii PLY files whose vertex element holds float properties (any order; f_rest_* are
    taken as the channel-major higher bands of the 3DGS convention)."""
    raw = Path(path).read_bytes()
    end = raw.index(b"end_header\n") + len(b"end_header\n")
    lines = raw[:end].decode("ascii").splitlines()
    if lines[0].strip() != "ply":
        raise ValueError(f"{path}: not a PLY file")
    fmt, count, props, in_vertex = None, None, [], False
    for ln in lines[1:]:
        t = ln.split()
        if not t:
            continue
        if t[0] == "format":
            fmt = t[1]
        elif t[0] == "element":
            in_vertex = t[1] == "vertex"
            if in_vertex:
                count = int(t[2])
        elif t[0] == "property" and in_vertex:
            if t[1] not in ("float", "float32"):
                raise ValueError(f"{path}: vertex property {t[-1]} is {t[1]}, expected float")
            props.append(t[-1])
    if count is None or fmt not in ("binary_little_endian", "ascii"):
        raise ValueError(f"{path}: unsupported PLY (format {fmt})")
    if fmt == "ascii":
        table = np.array(raw[end:].decode("ascii").split()[: count * len(props)], dtype=np.float32).reshape(count, len(props))
    else:
        table = np.frombuffer(raw, dtype="<f4", count=count * len(props), offset=end).reshape(count, len(props))
    col = {n: i for i, n in enumerate(props)}
    pick = lambda names: torch.from_numpy(np.stack([table[:, col[n]] for n in names], axis=-1).astype(np.float32))
    rest = sorted((n for n in props if n.startswith("f_rest_")), key=lambda n: int(n.split("_")[-1]))
    dc = pick(["f_dc_0", "f_dc_1", "f_dc_2"])[..., None]                       # [g,3,1]
    if rest:
        hi = pick(rest).reshape(count, 3, len(rest) // 3)                       # channel-major, as 3DGS stores them
        dc = torch.cat([dc, hi], dim=-1)
    w, x, y, z = pick(["rot_0", "rot_1", "rot_2", "rot_3"]).unbind(-1)
    return dict(means=pick(["x", "y", "z"]), scales=pick(["scale_0", "scale_1", "scale_2"]).exp(), rotations=torch.stack([x, y, z, w], -1),
                harmonics=dc, opacities=torch.sigmoid(pick(["opacity"])[..., 0]))


def gaussians_from_ply(path: Path, sh_coeffs: int | None = None) -> Gaussians:
    """The decoder's input built from a ply: [1,g,...] tensors, covariances R S S^T R^T (gaussians.py:33-44), harmonics
    zero-padded (or cut) to ``sh_coeffs`` coefficients per channel."""
    d = load_ply(path)
    i, j, k, r = d["rotations"].unbind(-1)
    two_s = 2 / ((d["rotations"] ** 2).sum(-1) + 1e-8)
    Rm = torch.stack([1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                      two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                      two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)], -1).reshape(-1, 3, 3)
    S = torch.diag_embed(d["scales"])
    cov = Rm @ S @ S.transpose(-1, -2) @ Rm.transpose(-1, -2)
    sh = d["harmonics"]
    if sh_coeffs is not None and sh.shape[-1] != sh_coeffs:
        out = torch.zeros(sh.shape[0], 3, sh_coeffs)
        n = min(sh_coeffs, sh.shape[-1])
        out[..., :n] = sh[..., :n]
        sh = out
    return Gaussians(d["means"][None].contiguous(), cov[None].contiguous(), sh[None].contiguous(), d["opacities"][None].contiguous())
