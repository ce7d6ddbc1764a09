#!/usr/bin/env python
"""Generate tests/golden/final_head.npz by EXECUTING the reference's ``Final.call``
(/root/reference/models/enet/enet_modules.py:1359-1381) -- build container only.

The method is located with ``ast`` and compiled in memory (nothing is copied); ``tf`` is a NumPy stand-in
whose ``tf.nn.conv2d_transpose`` implements the documented TF semantics generically (gradient of conv2d
w.r.t. its input for any stride and SAME/VALID padding, scatter-add form, float64 accumulation rounded to
float32) and checks the ``output_shape`` it is handed.  This pins what the reference asks of TF (filter
layout [kh,kw,classes,in], strides [1,2,2,1], padding "SAME", output 2h x 2w); TF's kernel itself stays
unpinned (TensorFlow 1.13.2 is not installable here)."""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ALS_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "models", "enet", "enet_modules.py")


class _Scope:
    def __init__(self, *a, **k): pass
    def __enter__(self): return self
    def __exit__(self, *e): return False


class _Tensor(np.ndarray):
    """ndarray whose .shape has as_list(), like a tf.Tensor's static shape."""
    class _Shape(tuple):
        def as_list(self): return list(self)

    @property
    def shape(self):
        return _Tensor._Shape(np.ndarray.shape.__get__(self))


def conv2d_transpose(value, filter, output_shape, strides, padding="SAME", name=None):
    v = np.asarray(value, np.float64)
    k = np.asarray(filter, np.float64)
    B, h, w, cin = v.shape
    kh, kw, cout, cin2 = k.shape
    assert cin == cin2 and strides[0] == strides[3] == 1
    sy, sx = strides[1], strides[2]
    out_shape = [int(x) for x in np.asarray(output_shape).tolist()]
    H, W = out_shape[1], out_shape[2]
    assert out_shape[0] == B and out_shape[3] == cout
    if padding == "SAME":
        assert h == -(-H // sy) and w == -(-W // sx), "output_shape inconsistent with SAME padding"
        pad_y = max((h - 1) * sy + kh - H, 0); pad_x = max((w - 1) * sx + kw - W, 0)
    else:
        assert h == (H - kh) // sy + 1 and w == (W - kw) // sx + 1
        pad_y = pad_x = 0
    top, left = pad_y // 2, pad_x // 2
    out = np.zeros((B, H, W, cout), np.float64)
    for i in range(h):
        for ky in range(kh):
            y = i * sy + ky - top
            if not 0 <= y < H:
                continue
            for j in range(w):
                for kx in range(kw):
                    x = j * sx + kx - left
                    if 0 <= x < W:
                        out[:, y, x, :] += v[:, i, j, :] @ k[ky, kx].T
    return out.astype(np.float32)


def make_tf():
    tf = types.SimpleNamespace()
    tf.name_scope = _Scope
    tf.shape = lambda x: np.asarray(np.ndarray.shape.__get__(x))
    tf.stack = lambda xs: np.asarray([int(x) for x in xs])
    tf.nn = types.SimpleNamespace(conv2d_transpose=conv2d_transpose)
    return tf


def final_call_code():
    with open(SRC) as f:
        tree = ast.parse(f.read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == "Final":
            for fn in node.body:
                if isinstance(fn, ast.FunctionDef) and fn.name == "call":
                    return compile(ast.Module(body=[fn], type_ignores=[]), SRC, "exec"), (fn.lineno, fn.end_lineno)
    raise RuntimeError("Final.call not found")


def main():
    code, lines = final_call_code()
    g = {"tf": make_tf()}
    exec(code, g)
    call = g["call"]
    rng = np.random.default_rng(20191013)
    out = {"meta_reference_lines": np.asarray(lines, np.int64)}
    names = []
    for name, (B, h, w, C) in {"c19": (2, 5, 7, 19), "c6": (1, 4, 9, 6), "c66": (1, 3, 4, 66), "c2": (2, 2, 2, 2)}.items():
        feat = rng.standard_normal((B, h, w, 16)).astype(np.float32).view(_Tensor)
        kern = (0.4 * rng.standard_normal((3, 3, C, 16))).astype(np.float32)
        self = types.SimpleNamespace(_conv_scope="ConvTransposed/", dilation_rate=(1, 1), classes=C, kernel=kern,
                                     padding="SAME")
        logits = call(self, feat)
        assert logits.shape == (B, 2 * h, 2 * w, C) and logits.dtype == np.float32
        out[name + ".features"] = np.asarray(feat)
        out[name + ".kernel"] = kern
        out[name + ".logits"] = logits
        names.append(name)
    out["cases"] = np.asarray(names)
    path = os.path.join(HERE, "final_head.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "Final.call lines", lines)


if __name__ == "__main__":
    main()
