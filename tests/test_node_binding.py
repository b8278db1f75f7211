"""bindings/node: the N-API shim and TypeScript surface a maintainer drops into
node-speex-resampler. Node is not in this image, so on the CPU we check what
can be checked without it: addon.c type-checks against a declaration-only <node_api.h>, links
against libspeexb200.so with nothing unresolved except N-API itself, binds only symbols that
include/speexb200.h declares, and index.ts carries the reference's fixed error texts."""
import os
import re
import shutil
import subprocess

import pytest

from node_speex_resampler_b200._lib import DECLARED_SYMBOLS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NODE = os.path.join(ROOT, "bindings", "node")
ADDON = os.path.join(NODE, "src", "addon.c")
INC = ["-I" + os.path.join(NODE, "test", "stub"), "-I" + os.path.join(ROOT, "include")]

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")


def test_addon_type_checks_against_napi_declarations():
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", *INC, ADDON], check=True)


def test_addon_links_and_binds_only_declared_symbols(tmp_path):
    so = tmp_path / "speexb200_node.so"
    libdir = os.path.join(ROOT, "node_speex_resampler_b200")
    subprocess.run(["gcc", "-std=c99", "-O1", "-shared", "-fPIC", *INC, ADDON, "-o", str(so),
                    "-L" + libdir, "-lspeexb200", "-Wl,-rpath," + libdir], check=True)
    undefined = subprocess.run(["nm", "-D", "--undefined-only", str(so)], check=True, capture_output=True,
                               text=True).stdout.split("\n")
    names = {ln.split()[-1].split("@")[0] for ln in undefined if ln.strip()}
    ours = {n for n in names if n.startswith(("speex_", "spxb_"))}
    assert ours, "the addon must call into libspeexb200"
    assert ours <= set(DECLARED_SYMBOLS), ours - set(DECLARED_SYMBOLS)
    # the five symbols the reference binds (src/index.ts:6-16) minus get_rate, which it never calls
    assert {"speex_resampler_init", "speex_resampler_destroy", "speex_resampler_process_interleaved_int",
            "speex_resampler_strerror"} <= ours
    napi = {n for n in names if n.startswith("napi_")}
    assert napi, "N-API entry points stay unresolved until Node loads the module"
    exported = subprocess.run(["nm", "-D", "--defined-only", str(so)], check=True, capture_output=True,
                              text=True).stdout
    assert "napi_register_module_v1" in exported


def test_typescript_surface_matches_the_reference_contract():
    ts = open(os.path.join(NODE, "src", "index.ts")).read()
    # the two fixed pre-check messages of src/index.ts:52,56 and the public surface
    assert "You need to wait for SpeexResampler.initPromise before calling this method" in ts
    assert "Chunk length should be a multiple of channels * 2 bytes" in ts
    for needle in ("static initPromise", "processChunk(chunk: Buffer): Buffer", "static processChunks(",
                   "export class SpeexResamplerTransform extends Transform", "export default SpeexResampler",
                   "export class SpeexResamplerBatchTransform extends Transform"):
        assert needle in ts, needle
    # the capacity rule of src/index.ts:80-95 is the same expression the Python mirror uses
    assert re.search(r"Math\.ceil\(bytes \* this\.outRate / this\.inRate\)", ts)
    assert re.search(r"Math\.trunc\(this\._outBufferSize / this\.channels / BYTES_PER_SAMPLE\)", ts)
    py = open(os.path.join(ROOT, "node_speex_resampler_b200", "resampler.py")).read()
    assert "math.ceil(nbytes * self.outRate / self.inRate)" in py


def test_addon_selftest_builds_and_fails_loudly_without_a_gpu():
    """the addon linked with the in-process N-API stand-in (bindings/node/test/fake_napi.c); the
    GPU suite runs it for real (tests/test_parity_gpu.py::test_node_addon_executes_like_the_mirror)"""
    import __graft_entry__ as G
    exe = G.build_node_addon_selftest()
    import node_speex_resampler_b200 as pkg
    if pkg.lib().spxb_device_count() <= 0:
        r = subprocess.run([exe], capture_output=True, text=True)
        assert r.returncode == 3 and "no CUDA device" in r.stdout
