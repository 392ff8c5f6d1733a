"""Paraview VTI output of fields that live on the GPU (second half).

File format and naming follow the reference byte for byte (``VoxelDomain.write_to_vti``,
pymoto/common/domain.py:452-604; ``WriteToVTI``, pymoto/modules/io.py:286-348): ImageData, ``header_type="UInt64"``,
one base64 block per ``DataArray`` preceded by the base64 of its encoded length, Float32 payload, vectors whose size is a
multiple of ``nel`` are cell data, multiples of ``nnodes`` point data, 2-component nodal vectors of a 2-D domain are
padded to 3 components, block vectors get ``(i)`` suffixes, complex vectors are split into ``(real)`` / ``(imag)``.

What is new: a vector may be a CUDA tensor.  It is converted FP64 -> FP32 (and padded 2 -> 3 components) by a libpmb
kernel (``pmb_pack_f32``) so only the 4-byte payload crosses PCIe; host arrays take numpy's ``astype`` as in the
reference.  Both routes round to nearest-even, so the files are identical.
"""
import base64
import os
import struct
import sys
import warnings
from pathlib import Path

import numpy as np

from . import _lib
from . import device as dv
from .core import Module


def _to_f32(vec, ncomp_in=1, ncomp_out=1):
    """Float32 payload (host bytes-like) of one real 1-D vector; components padded with zeros from ncomp_in to ncomp_out."""
    if dv.is_device(vec):
        import torch

        src = vec.reshape(-1)
        if src.dtype != torch.float64:
            src = src.to(torch.float64)
        src = src.contiguous()
        nitems = src.numel() // ncomp_in
        out = torch.empty(nitems * ncomp_out, dtype=torch.float32, device=src.device)
        _lib.call("pmb_pack_f32", nitems, ncomp_in, ncomp_out, dv.ptr(src), dv.ptr(out), dv.stream())
        return out.cpu().numpy()
    v = np.asarray(vec).astype(np.float32)
    if ncomp_out != ncomp_in:
        pad = np.zeros((v.size // ncomp_in) * ncomp_out, dtype=np.float32)
        for c in range(ncomp_in):
            pad[c::ncomp_out] = v[c::ncomp_in]
        v = pad
    return v


def _is_complex(vec):
    if dv.is_device(vec):
        return vec.is_complex()
    return np.iscomplexobj(vec)


def _size(vec):
    return int(vec.numel()) if dv.is_device(vec) else int(np.asarray(vec).size)


def _split(vec, unit):
    """(ncomponents, [sub-vectors]) of a 1-D vector or a 2-D block vector whose axis ``vecax`` is a multiple of ``unit``."""
    shape = tuple(vec.shape)
    assert len(shape) <= 2, "Only for 1D and 2D numpy arrays"
    vecax = next((i for i, s in enumerate(shape) if s % unit == 0), None)
    if vecax is None:  # size is a multiple of the unit but no single axis is (the reference fails here as well)
        raise ValueError(f"no axis of shape {shape} is a multiple of {unit}")
    ncomp = shape[vecax] // unit
    if len(shape) == 1:
        return ncomp, [vec]
    other = (vecax + 1) % 2
    subs = [(vec[:, i] if other == 1 else vec[i, :]) for i in range(shape[other])]
    return ncomp, subs


def _write_array(file, name, ncomp, payload, len_enc):
    file.write(f'<DataArray type="Float32" Name="{name}" NumberOfComponents="{ncomp}" format="binary">\n'.encode())
    enc = base64.b64encode(payload)
    file.write(base64.b64encode(struct.pack(len_enc, len(enc))))
    file.write(enc)
    file.write(b"\n</DataArray>\n")


def write_to_vti(domain, vectors: dict, filename="out.vti", scale=1.0):
    """Write numpy arrays and / or CUDA tensors to a VTI file (see the module docstring)."""
    ext = ".vti"
    if ext not in os.path.splitext(filename)[-1].lower():
        filename += ext
    nelx, nely, nelz = int(domain.nelx), int(domain.nely), int(getattr(domain, "nelz", 0) or 0)
    nel, nnodes = int(domain.nel), int(domain.nnodes)
    dim = 2 if nelz == 0 else 3
    point_dat, cell_dat = {}, {}
    for key, vec in vectors.items():
        if _size(vec) % nel == 0:
            cell_dat[key] = vec
        elif _size(vec) % nnodes == 0:
            point_dat[key] = vec
        else:
            warnings.warn(f"Vector {key} is neither cell- nor point-data. Skipping vector...")
    if len(point_dat) == 0 and len(cell_dat) == 0:
        warnings.warn(f"Nothing to write to {filename}. Skipping file...")
        return
    len_enc = ("<" if sys.byteorder == "little" else ">") + "Q"
    origin = np.asarray(getattr(domain, "origin", np.zeros(3)), dtype=float)
    h = np.asarray(domain.element_size, dtype=float)
    with open(filename, "wb") as file:
        file.write(b'<?xml version="1.0"?>\n')
        byte_order = "LittleEndian" if sys.byteorder == "little" else "BigEndian"
        file.write(f'<VTKFile type="ImageData" version="0.1" header_type="UInt64" byte_order="{byte_order}">\n'.encode())
        file.write(f'<ImageData WholeExtent="0 {nelx} 0 {nely} 0 {nelz}"'.encode())
        file.write(f' Origin="{origin[0] * scale} {origin[1] * scale} {origin[2] * scale}"'.encode())
        dx, dy, dz = h[0:3] * scale
        file.write(f' Spacing="{dx} {dy} {dz}">\n'.encode())
        file.write(f'<Piece Extent="0 {nelx} 0 {nely} 0 {nelz}">\n'.encode())

        def section(tag, items, unit, point):
            file.write(f"<{tag}>\n".encode())
            for key, vec in items.items():
                ncomp, subs = _split(vec, unit)
                pad = point and ncomp == 2 and dim == 2  # enables Paraview's deform button on 2-D displacement fields
                nzeros = int(np.ceil(np.log10(len(subs)))) if len(subs) > 1 else 0
                for i, sub in enumerate(subs):
                    name = key
                    if len(subs) > 1:
                        name += f"({i:0{nzeros}d})" if point else f"({i})"
                    if _is_complex(sub):
                        parts = [(sub.real, name + "(real)"), (sub.imag, name + "(imag)")]
                    else:
                        parts = [(sub, name)]
                    for v, t in parts:
                        ncout = 3 if pad else ncomp
                        _write_array(file, t, ncout, _to_f32(v, ncomp, ncout), len_enc)
            file.write(f"</{tag}>\n".encode())

        if point_dat:
            section("PointData", point_dat, nnodes, True)
        if cell_dat:
            section("CellData", cell_dat, nel, False)
        file.write(b"</Piece>\n")
        file.write(b"</ImageData>\n")
        file.write(b"</VTKFile>")


class WriteToVTI(Module):
    """Module form (pymoto/modules/io.py:286-348): writes its input states every ``interval`` calls; the signal tags name
    the arrays; ``overwrite=False`` numbers the files ``name.0000.vti`` ... by iteration."""

    def __init__(self, domain, saveto: str, overwrite: bool = False, scale=1.0, interval=1):
        self.domain = domain
        self.saveto = saveto
        Path(saveto).parent.mkdir(parents=True, exist_ok=True)
        ext = os.path.splitext(saveto)[-1].lower()
        supported_ext = (".vti",)
        if ext not in supported_ext:
            raise ValueError(f"Extension `{ext}` is not supported. Supported extensions are {supported_ext} ")
        self.iter = 0
        self.scale = scale
        self.overwrite = overwrite
        self.interval = interval

    def __call__(self, *args):
        if self.iter % self.interval != 0:
            self.iter += 1
            return
        if len(args) == 0:
            raise ValueError("Nothing to write to VTI file")
        data = {}
        sig_in = getattr(self, "sig_in", None)
        for i, x in enumerate(args):
            nam = f"inp{i:d}"
            if sig_in is not None:
                nam = getattr(sig_in[i], "tag", nam)
            data[nam] = x
        pth = os.path.splitext(self.saveto)
        filen = pth[0] + pth[1] if self.overwrite else pth[0] + ".{0:04d}".format(self.iter) + pth[1]
        write_to_vti(self.domain, data, filename=filen, scale=self.scale)
        self.iter += 1
