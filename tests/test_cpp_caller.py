"""The C ABI from a C++ program with no Python / torch in the process: examples/cpp/dn_cpp_smoke
dlopen()s the library, reads a serialised engine file (demonet_b200/export.py) and runs the two images of the
reference's libtorch test (test/tracing/test_demonet_tracing.cpp:31-33: rand{3,320,320} and rand{3,256,275})."""
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

import demonet_b200
from demonet_b200 import _C, export
from oracle import weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP_DIR = os.path.join(ROOT, "examples", "cpp")
EXE = os.path.join(CPP_DIR, "dn_cpp_smoke")


def _build():
    subprocess.check_call(["make", "-C", CPP_DIR, "-s"])
    assert os.path.exists(EXE)


def test_cpp_program_builds_and_binds_every_symbol():
    """No GPU needed: the program compiles against include/demonet_b200.h and resolves its entry points in both builds."""
    _build()
    for dt in _C.ACT_DTYPES:
        out = subprocess.run([EXE, "--symbols", _C.LIB_PATHS[dt]], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr


def test_engine_file_layout(tmp_path):
    """The serialised engine file carries the struct sizes of the ctypes binding = those of the C header the program uses."""
    m = demonet_b200.ssdlite320_mobilenet_v3_large()
    info = export.save_engine_file(m, str(tmp_path / "v3.dneng"))
    raw = open(tmp_path / "v3.dneng", "rb").read()
    assert raw[:8] == export.MAGIC
    abi, dt, sz_desc, sz_op, sz_buf, P = struct.unpack("<6i", raw[8:32])
    (blob_bytes,) = struct.unpack("<q", raw[32:40])
    assert (abi, P) == (_C.ABI_VERSION, 3234) and dt == (1 if m.act_dtype == "fp16" else 0)
    assert len(raw) == 40 + sz_desc + info["n_ops"] * sz_op + info["n_bufs"] * sz_buf + P * 16 + blob_bytes


@pytest.mark.gpu
@pytest.mark.parametrize("act_dtype", ["fp16", "bf16"])
def test_cpp_program_reproduces_the_python_detections(tmp_path, act_dtype):
    _build()
    model = demonet_b200.ssdlite320_mobilenet_v3_large(act_dtype=act_dtype)
    model.load_state_dict(weights.seeded_state_dict(model.state_dict()))
    g = torch.Generator().manual_seed(3)
    img0, img1 = torch.rand(3, 320, 320, generator=g), torch.rand(3, 256, 275, generator=g)
    eng_file, img_file, out_file = (str(tmp_path / n) for n in ("v3.dneng", "images.f32", "dets.bin"))
    export.save_engine_file(model, eng_file)
    with open(img_file, "wb") as f:
        f.write(img0.numpy().tobytes())
        f.write(img1.numpy().tobytes())
    run = subprocess.run([EXE, _C.LIB_PATHS[act_dtype], eng_file, img_file, out_file], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "graph replays" in run.stdout
    raw = open(out_file, "rb").read()
    B, D = struct.unpack("<2i", raw[:8])
    assert (B, D) == (2, 300)
    off = 8
    counts = np.frombuffer(raw, np.int32, B, off); off += 4 * B
    boxes = np.frombuffer(raw, np.float32, B * D * 4, off).reshape(B, D, 4); off += 16 * B * D
    scores = np.frombuffer(raw, np.float32, B * D, off).reshape(B, D); off += 4 * B * D
    labels = np.frombuffer(raw, np.int64, B * D, off).reshape(B, D)
    want = model.cuda()([img0.cuda(), img1.cuda()])
    for i, w in enumerate(want):
        n = int(counts[i])
        assert n == w["scores"].numel()
        assert np.array_equal(scores[i, :n], w["scores"].cpu().numpy()) and np.array_equal(labels[i, :n], w["labels"].cpu().numpy())
        assert np.array_equal(boxes[i, :n], w["boxes"].cpu().numpy())
    assert float(boxes[1, :, 0::2].max()) <= 275.0 + 1e-3 and float(boxes[1, :, 1::2].max()) <= 256.0 + 1e-3
