"""Generates the golden fixtures in this directory from the REFERENCE ITSELF.

Runs oracle/_ref/OpticalFlow_ref (the unmodified reference sources compiled by oracle/ref/build_ref.sh;
needs /root/reference, so this script only works in the build container) with --tap on two small
seeded inputs and stores inputs + selected reference outputs as compressed .npz:

  sphere3_vertex.npz   --in A.ply B.ply --out r.ply on the 258-vertex octahedron sphere, 10 iterations
  torus_texture.npz    --mesh m.ply --in A.png B.png --out r.png --eLength 0.08 on a 24x12 uv torus, 48x48 texels
  sample_texture_tool.npz  oracle/_ref/SampleTextureToVertices_ref on the uv torus (ascii, subdivided, binary): files in, files out
  sphere3_modes.npz    the 258-vertex sphere again with 4 iterations of --vfMode 1, --vfMode 2 --cMode 0|1|2 (taps) and of
                       --dogWeight 0.5 (the 6-channel blend: output colours only, the tap build is 3-channel) and of --log (comparison signals, last flow, advected and output colours)

  sphere7_vertex.npz / sphere8_vertex.npz   the 65 538- and 262 146-vertex spheres (BASELINE.json configs[2]'s generator at the two
                       sizes below the headline one that the reference finishes in minutes), 3 iterations: per-iteration flow on
                       every 16th / 64th triangle with its global norm, advected colours on every 16th / 64th vertex, all output colours
  example.npz          the reference's own Example/ (BASELINE.json configs[0] and [1]): mesh.ply, A.png, B.png byte for byte (input
                       vectors the reference holds), the reference's output picture for --mesh mesh.ply --in A.png B.png and its output
                       colours for --in A.ply B.ply (A.ply / B.ply = the reference's SampleTextureToVertices_ref --eLength 0.006)

    python tests/golden/make_golden.py [midsize 7|8 | example | modes | tool]
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from meshopticalflow_b200 import synthetic  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "OpticalFlow_ref")


def load_taps(d):
    return {f[:-4]: np.load(os.path.join(d, f)) for f in os.listdir(d) if f.endswith(".npy")}


def keep(taps, names, iterations=10, per_iter=("tFlowField",)):
    out = {n: taps[n] for n in names}
    for i in range(iterations):
        for n in per_iter:
            out["it%02d.%s" % (i, n)] = taps["it%02d.%s" % (i, n)]
    return out


def sphere():
    from PIL import Image  # noqa: F401  (only to fail early if the harness image lacks it)
    v, t = synthetic.octahedron_sphere(3)
    a, b = synthetic.smooth_rgb_pair(v, 0)
    with tempfile.TemporaryDirectory() as d:
        synthetic.write_ply_colored(os.path.join(d, "A.ply"), v, a, t)
        synthetic.write_ply_colored(os.path.join(d, "B.ply"), v, b, t)
        subprocess.check_call([REF, "--in", "A.ply", "B.ply", "--out", "r.ply", "--tap", "tap"], cwd=d, stdout=subprocess.DEVNULL)
        taps = load_taps(os.path.join(d, "tap"))
        out = synthetic.read_ply(os.path.join(d, "r.ply"))
    data = keep(taps, ["vertices", "triangles", "g", "oppositeEdge", "xform_linear", "xform_constant", "reducedEdgeIndex", "expandedEdgeIndex",
                       "positiveOrientedEdge", "signals0", "signals1", "advected0", "advected1", "sMass.rowptr", "sMass.col", "sMass.val",
                       "sStiffness.val", "smoothOperator.rowptr", "smoothOperator.col", "smoothOperator.val", "prolongation.col", "prolongation.val"],
                per_iter=("tFlowField", "x", "smoothed0", "resampled1", "dataTerm", "rhs"))
    data["input_vertices_f32"] = v.astype(np.float32)
    data["input_a"], data["input_b"] = a, b
    data["output_rgb"] = np.stack([out["vertex"][k] for k in ("red", "green", "blue")], 1).astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "sphere3_vertex.npz"), **data)
    print("sphere3_vertex.npz", os.path.getsize(os.path.join(HERE, "sphere3_vertex.npz")) // 1024, "KiB")


def torus():
    from PIL import Image
    v, t, uv = synthetic.uv_torus(24, 12)
    ta, tb = synthetic.smooth_texture_pair(48, 48, 1)
    with tempfile.TemporaryDirectory() as d:
        synthetic.write_ply_textured(os.path.join(d, "m.ply"), v, t, uv)
        Image.fromarray(ta).save(os.path.join(d, "A.png"))
        Image.fromarray(np.dstack([tb, np.full(tb.shape[:2], 255, np.uint8)])).save(os.path.join(d, "B.png"))  # RGBA: alpha must be dropped
        subprocess.check_call([REF, "--mesh", "m.ply", "--in", "A.png", "B.png", "--out", "r.png", "--eLength", "0.08", "--tap", "tap"], cwd=d,
                              stdout=subprocess.DEVNULL)
        taps = load_taps(os.path.join(d, "tap"))
        pixels = np.asarray(Image.open(os.path.join(d, "r.png")))
        png_a = open(os.path.join(d, "A.png"), "rb").read()
        png_b = open(os.path.join(d, "B.png"), "rb").read()
    data = keep(taps, ["vertices", "triangles", "triangleTextures", "oppositeEdge", "signals0", "signals1", "textureSource_tIdx", "textureSource_p",
                       "advected0", "advected1", "texture0", "texture1"])
    data["input_vertices_f32"], data["input_triangles"], data["input_uv"] = v, t, uv
    data["input_tex_a"], data["input_tex_b"] = ta, tb
    data["png_a"], data["png_b"] = np.frombuffer(png_a, dtype=np.uint8), np.frombuffer(png_b, dtype=np.uint8)
    data["output_pixels"] = pixels
    np.savez_compressed(os.path.join(HERE, "torus_texture.npz"), **data)
    print("torus_texture.npz", os.path.getsize(os.path.join(HERE, "torus_texture.npz")) // 1024, "KiB")


def tool():
    """SampleTextureToVertices (the sibling tool) on the uv torus: input files and the reference's output files, byte for byte."""
    from PIL import Image
    ref_tool = os.path.join(ROOT, "oracle", "_ref", "SampleTextureToVertices_ref")
    v, t, uv = synthetic.uv_torus(24, 12)
    ta, _ = synthetic.smooth_texture_pair(48, 48, 1)
    data = {}
    with tempfile.TemporaryDirectory() as d:
        synthetic.write_ply_textured(os.path.join(d, "m.ply"), v, t, uv)
        with open(os.path.join(d, "mb.ply"), "wb") as fp:  # the same mesh as binary records
            fp.write((f"ply\nformat binary_little_endian 1.0\nelement vertex {len(v)}\nproperty float x\nproperty float y\nproperty float z\n"
                      f"element face {len(t)}\nproperty list uchar int vertex_indices\nproperty list uchar float texcoord\nend_header\n").encode())
            fp.write(np.asarray(v, dtype="<f4").tobytes())
            rec = np.zeros(len(t), dtype=[("n", "u1"), ("i", "<i4", 3), ("m", "u1"), ("uv", "<f4", 6)])
            rec["n"], rec["i"], rec["m"], rec["uv"] = 3, t, 6, uv
            fp.write(rec.tobytes())
        Image.fromarray(ta).save(os.path.join(d, "A.png"))
        runs = {"plain": ["--in", "m.ply"], "subdivided": ["--in", "m.ply", "--eLength", "0.08"], "binary": ["--in", "mb.ply", "--eLength", "0.1"]}
        for name, flags in runs.items():
            subprocess.check_call([ref_tool, "--texture", "A.png", "--out", name + ".ply"] + flags, cwd=d, stdout=subprocess.DEVNULL)
            data["out_" + name] = np.frombuffer(open(os.path.join(d, name + ".ply"), "rb").read(), dtype=np.uint8)
        for f in ("m.ply", "mb.ply", "A.png"):
            data["in_" + f] = np.frombuffer(open(os.path.join(d, f), "rb").read(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "sample_texture_tool.npz"), **data)
    print("sample_texture_tool.npz", os.path.getsize(os.path.join(HERE, "sample_texture_tool.npz")) // 1024, "KiB")


MODES = {"conformal": ["--vfMode", "1"], "connection0": ["--vfMode", "2"], "connection1": ["--vfMode", "2", "--cMode", "1"],
         "connection2": ["--vfMode", "2", "--cMode", "2"]}


def modes():
    v, t = synthetic.octahedron_sphere(3)
    a, b = synthetic.smooth_rgb_pair(v, 0)
    data = {"input_vertices_f32": v.astype(np.float32), "triangles": t.astype(np.int32), "input_a": a, "input_b": b}
    with tempfile.TemporaryDirectory() as d:
        synthetic.write_ply_colored(os.path.join(d, "A.ply"), v, a, t)
        synthetic.write_ply_colored(os.path.join(d, "B.ply"), v, b, t)
        for name, flags in MODES.items():
            subprocess.check_call([REF, "--in", "A.ply", "B.ply", "--out", name + ".ply", "--iterations", "4", "--tap", "tap_" + name] + flags, cwd=d,
                                  stdout=subprocess.DEVNULL)
            taps = load_taps(os.path.join(d, "tap_" + name))
            for k in ("smoothOperator.rowptr", "smoothOperator.col", "smoothOperator.val", "advected0", "advected1"):
                data[name + "." + k] = taps[k]
            for i in range(4):
                data["%s.it%02d.tFlowField" % (name, i)] = taps["it%02d.tFlowField" % i]
                data["%s.it%02d.coeffs" % (name, i)] = taps["it%02d.coeffs" % i]
            out = synthetic.read_ply(os.path.join(d, name + ".ply"))
            data[name + ".output_rgb"] = np.stack([out["vertex"][k] for k in ("red", "green", "blue")], 1).astype(np.uint8)
        subprocess.check_call([REF, "--in", "A.ply", "B.ply", "--out", "blend.ply", "--iterations", "4", "--dogWeight", "0.5"], cwd=d, stdout=subprocess.DEVNULL)
        out = synthetic.read_ply(os.path.join(d, "blend.ply"))
        data["blend.output_rgb"] = np.stack([out["vertex"][k] for k in ("red", "green", "blue")], 1).astype(np.uint8)
        subprocess.check_call([REF, "--in", "A.ply", "B.ply", "--out", "log.ply", "--iterations", "4", "--log", "--tap", "tap_log"], cwd=d, stdout=subprocess.DEVNULL)
        out = synthetic.read_ply(os.path.join(d, "log.ply"))
        data["log.output_rgb"] = np.stack([out["vertex"][k] for k in ("red", "green", "blue")], 1).astype(np.uint8)
        taps = load_taps(os.path.join(d, "tap_log"))
        data["log.signals0"], data["log.advected0"], data["log.it03.tFlowField"] = taps["signals0"], taps["advected0"], taps["it03.tFlowField"]
    np.savez_compressed(os.path.join(HERE, "sphere3_modes.npz"), **data)
    print("sphere3_modes.npz", os.path.getsize(os.path.join(HERE, "sphere3_modes.npz")) // 1024, "KiB")


def midsize(level):
    """Per-vertex alignment of the synthetic sphere of `level` (float32 positions, like any PLY), 3 iterations, taps subsampled."""
    stride = 16 if level <= 7 else 64
    v, t = synthetic.octahedron_sphere(level)
    a, b = synthetic.smooth_rgb_pair(v, 0)
    with tempfile.TemporaryDirectory() as d:
        synthetic.write_ply_colored(os.path.join(d, "A.ply"), v, a, t)
        synthetic.write_ply_colored(os.path.join(d, "B.ply"), v, b, t)
        subprocess.check_call([REF, "--in", "A.ply", "B.ply", "--out", "r.ply", "--iterations", "3", "--tap", "tap"], cwd=d, stdout=subprocess.DEVNULL)
        tapdir = os.path.join(d, "tap")
        data = {"level": np.int32(level), "stride": np.int32(stride), "iterations": np.int32(3)}
        for i in range(3):
            f = np.load(os.path.join(tapdir, "it%02d.tFlowField.npy" % i))
            data["it%02d.tFlowField.sub" % i] = f[::stride].copy()
            data["it%02d.tFlowField.norm" % i] = np.float64(np.linalg.norm(f))
            data["it%02d.x.norm" % i] = np.float64(np.linalg.norm(np.load(os.path.join(tapdir, "it%02d.x.npy" % i))))
        for k in ("advected0", "advected1"):
            data[k + ".sub"] = np.load(os.path.join(tapdir, k + ".npy"))[::stride].copy()
        opp = np.load(os.path.join(tapdir, "oppositeEdge.npy"))
        red = np.load(os.path.join(tapdir, "reducedEdgeIndex.npy"))
        data["oppositeEdge.sub"], data["reducedEdgeIndex.sub"] = opp[::stride].copy(), red[::stride].copy()
        data["oppositeEdge.sum"], data["reducedEdgeIndex.sum"] = np.int64(opp.astype(np.int64).sum()), np.int64(red.astype(np.int64).sum())
        out = synthetic.read_ply(os.path.join(d, "r.ply"))
    data["output_rgb"] = np.stack([out["vertex"][k] for k in ("red", "green", "blue")], 1).astype(np.uint8)
    name = "sphere%d_vertex.npz" % level
    np.savez_compressed(os.path.join(HERE, name), **data)
    print(name, os.path.getsize(os.path.join(HERE, name)) // 1024, "KiB")


def example():
    """The reference's Example/ through the reference binary: the texture configuration and the per-vertex configuration."""
    from PIL import Image
    ref_tool = os.path.join(ROOT, "oracle", "_ref", "SampleTextureToVertices_ref")
    src = os.path.join(os.environ.get("MOF_REFERENCE", "/root/reference"), "Example")
    data = {}
    for f in ("mesh.ply", "A.png", "B.png"):
        data["in_" + f] = np.frombuffer(open(os.path.join(src, f), "rb").read(), dtype=np.uint8)
    with tempfile.TemporaryDirectory() as d:
        for f in ("mesh.ply", "A.png", "B.png"):
            open(os.path.join(d, f), "wb").write(data["in_" + f].tobytes())
        subprocess.check_call([REF, "--mesh", "mesh.ply", "--in", "A.png", "B.png", "--out", "result.png"], cwd=d, stdout=subprocess.DEVNULL)
        data["texture_output_pixels"] = np.asarray(Image.open(os.path.join(d, "result.png")))
        for n in ("A", "B"):
            subprocess.check_call([ref_tool, "--in", "mesh.ply", "--texture", n + ".png", "--out", n + ".ply", "--eLength", "0.006"], cwd=d, stdout=subprocess.DEVNULL)
        subprocess.check_call([REF, "--in", "A.ply", "B.ply", "--out", "result.ply"], cwd=d, stdout=subprocess.DEVNULL)
        out = synthetic.read_ply(os.path.join(d, "result.ply"))
        data["vertex_output_rgb"] = np.stack([out["vertex"][k] for k in ("red", "green", "blue")], 1).astype(np.uint8)
        data["vertex_output_xyz"] = np.stack([out["vertex"][k] for k in ("x", "y", "z")], 1).astype(np.float32)
        data["vertex_output_faces_crc"] = np.int64(np.asarray(out["face"]["vertex_indices"], dtype=np.int64).sum())
    np.savez_compressed(os.path.join(HERE, "example.npz"), **data)
    print("example.npz", os.path.getsize(os.path.join(HERE, "example.npz")) // 1024, "KiB")


if __name__ == "__main__":
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/OpticalFlow_ref is missing: run oracle/ref/build_ref.sh (needs /root/reference)")
    if len(sys.argv) > 1 and sys.argv[1] in ("modes", "tool", "example"):
        {"modes": modes, "tool": tool, "example": example}[sys.argv[1]]()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[1] == "midsize":
        midsize(int(sys.argv[2]))
        sys.exit(0)
    sphere()
    torus()
    modes()
    tool()
