#!/usr/bin/env python3
"""Extract the golden vectors held by go-dsp's own test files into JSON.

Run in the build container (the reference tree is not present on the GPU box):
    python tests/golden/extract_reference_vectors.py [/root/reference]
It parses the Go table literals (data only -- no reference code is copied) and
writes tests/golden/reference_vectors.json, recording the file:line range each
table came from.  Complex values are stored as [re, im] pairs.
"""
import ast
import json
import math
import operator
import os
import re
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")


def go_block(path, var):
    """Return (python_value, 'file:first-last') of `var <name> = []T{ ... }`."""
    src = open(os.path.join(REF, path)).read().split("\n")
    start = next(i for i, l in enumerate(src) if re.match(r"\s*(var\s+)?%s\s*:?=\s*\[\]" % re.escape(var), l))
    depth, text, end = 0, [], start
    for i in range(start, len(src)):
        line = re.sub(r"//.*$", "", src[i])
        text.append(line)
        depth += line.count("{") - line.count("}")
        if depth == 0 and "{" in "".join(text):
            end = i
            break
    body = "\n".join(text)
    body = body[body.index("{"):]
    body = re.sub(r"(\[\])+\s*\w+\s*\{", "{", body)        # []float64{ / [][]complex128{ / []fftTest{
    body = re.sub(r"&\w+\{\}", "None", body)               # &PwelchOptions{}
    body = body.replace("{", "[").replace("}", "]")
    return literal(ast.parse(body.strip(), mode="eval").body), "%s:%d-%d" % (path, start + 1, end + 1)


NAMES = {"sqrt2_2": math.sqrt(2) / 2, "None": None}
BINOPS = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv}


def literal(node):
    """The reference tree is untrusted text: only numbers, lists, +-*/ of those, complex(a, b) and two known names are
    evaluated (no eval, no attribute access, no other calls)."""
    if isinstance(node, ast.Constant) and (node.value is None or isinstance(node.value, (int, float))):
        return node.value
    if isinstance(node, (ast.List, ast.Tuple)):
        return [literal(e) for e in node.elts]
    if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
        v = literal(node.operand)
        return -v if isinstance(node.op, ast.USub) else v
    if isinstance(node, ast.BinOp) and type(node.op) in BINOPS:
        return BINOPS[type(node.op)](literal(node.left), literal(node.right))
    if isinstance(node, ast.Name) and node.id in NAMES:
        return NAMES[node.id]
    if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id == "complex" and len(node.args) == 2 and not node.keywords:
        return complex(literal(node.args[0]), literal(node.args[1]))
    raise ValueError("unsupported construct in a reference table: %s" % ast.dump(node)[:80])


def enc(v):
    if isinstance(v, complex):
        return [v.real, v.imag]
    if isinstance(v, (list, tuple)):
        return [enc(e) for e in v]
    return v


def cplx_list(v):
    return [[complex(e).real, complex(e).imag] for e in v]


def wav_golden():
    """wav/wav_test.go:66-98 TestWav: the header fields, Samples and Duration wav.New must report for the two fixtures.
    The fixtures themselves are test data, copied next to this script: small.wav whole (84 KB) and float.wav cut after
    its first 16384 samples (the header still announces the full data size, which is what the test checks)."""
    src = open(os.path.join(REF, "wav/wav_test.go")).read()
    cases = {}
    for name in ("small.wav", "float.wav"):
        blk = src[src.index('"%s": {' % name):]
        blk = blk[: blk.index("typ:")]
        f = {k: literal(ast.parse(re.search(r"%s:\s*([0-9 /*]+)," % k, blk).group(1).strip(), mode="eval").body)
             for k in ("AudioFormat", "NumChannels", "SampleRate", "ByteRate", "BlockAlign", "BitsPerSample", "Samples", "Duration")}
        f["Samples"] = int(f["Samples"])
        cases[name] = f
    here = os.path.dirname(OUT)
    shutil.copyfile(os.path.join(REF, "wav/small.wav"), os.path.join(here, "small.wav"))
    with open(os.path.join(REF, "wav/float.wav"), "rb") as fsrc, open(os.path.join(here, "float_head.wav"), "wb") as fdst:
        fdst.write(fsrc.read(44 + 4 * 16384))
    os.chmod(os.path.join(here, "small.wav"), 0o644)
    return {"source": "wav/wav_test.go:66-98", "files": {"small.wav": "small.wav", "float.wav": "float_head.wav"}, "cases": cases,
            "note": "decoded sample values are not pinned by any reference test: ReadFloats (wav/wav.go:138-161) is pinned by its source lines only"}


def main():
    out = {"_generator": "tests/golden/extract_reference_vectors.py", "_reference": "maddyblue/go-dsp",
           "_tolerance": "dsputils/compare.go:23-25,94-96: |a-b| <= 1e-8 or |1-a/b| <= 1e-8 per real/imag part"}

    v, src = go_block("fft/fft_test.go", "fftTests")
    out["fft"] = {"source": src, "cases": [{"in": c[0], "out": cplx_list(c[1])} for c in v]}
    v, src = go_block("fft/fft_test.go", "fft2Tests")
    out["fft2"] = {"source": src, "cases": [{"in": c[0], "out": [cplx_list(r) for r in c[1]]} for c in v]}
    v, src = go_block("fft/fft_test.go", "fftnTests")
    out["fftn"] = {"source": src, "cases": [{"in": c[0], "dim": c[1], "out": cplx_list(c[2])} for c in v]}
    v, src = go_block("fft/fft_test.go", "reverseBitsTests")
    out["reverse_bits"] = {"source": src, "cases": [{"in": c[0], "sz": c[1], "out": c[2]} for c in v]}
    out["example_fft_real"] = {
        "source": "fft/fft_test.go:283-320",
        "note": "x(n) = sin(2*pi*n/8) + 0.5*sin(2*pi*n/4 + 3*pi/4), 8 samples; printed %.1f magnitude / phase(deg); "
                "phase forced to 0 when Float64Equal(magnitude, 0)",
        "mag": [0.0, 4.0, 2.0, 0.0, 0.0, 0.0, 2.0, 4.0], "phase_deg": [0.0, -90.0, 45.0, 0.0, 0.0, 0.0, -45.0, 90.0]}
    out["fft_multi"] = {"source": "fft/fft_test.go:251-259", "note": "N=256 ramp complex(i/N,0) must run"}

    v, src = go_block("spectral/pwelch_test.go", "pwelchTests")
    out["pwelch"] = {"source": src, "cases": [{"fs": c[0], "x": c[2], "p": c[3], "freqs": c[4]} for c in v]}
    v, src = go_block("spectral/spectral_test.go", "segmentTests")
    out["spectral_segment"] = {"source": src, "x": [1, 2, 3, 4, 5, 6, 7, 8, 9, 10],
                               "cases": [{"size": c[0], "noverlap": c[1], "out": c[2]} for c in v]}
    v, src = go_block("window/window_test.go", "windowTests")
    out["window"] = {"source": src, "cases": [{"L": c[0], "hamming": c[1], "hann": c[2], "bartlett": c[3],
                                               "flattop": c[4], "blackman": c[5]} for c in v]}
    v, src = go_block("dsputils/dsputils_test.go", "segmentTests")
    out["dsputils_segment"] = {"source": src, "n": 16, "cases": [{"segs": c[0], "noverlap": c[1], "slices": c[2]} for c in v]}
    out["matrix"] = {
        "source": "dsputils/matrix_test.go:23-46",
        "list": [1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 4, 3, 2, 1], "dims": [2, 3, 4],
        "dim_cases": [{"idx": [1, 0, -1], "out": [3, 4, 5, 6]}, {"idx": [0, -1, 2], "out": [3, 7, 1]},
                      {"idx": [-1, 1, 3], "out": [8, 0]}],
        "setdim": {"idx": [1, -1, 3], "values": [10, 11, 12]}}
    out["wav"] = wav_golden()
    with open(OUT, "w") as f:
        json.dump(enc(out), f, indent=1)
    print("wrote", OUT, {k: len(v["cases"]) for k, v in out.items() if isinstance(v, dict) and "cases" in v})


if __name__ == "__main__":
    main()
