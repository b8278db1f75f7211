"""SURVEY section 5 (sanitizers): the host-side planners under ASan + UBSan, and the oracle's C
restatement under the same on the golden inputs. The GPU kernels get compute-sanitizer
(scripts/gpu_sanitize.sh, profiles/sanitizer_r1.log)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "node_speex_resampler_b200", "csrc")
SAN = ["-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer", "-g", "-O1"]

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")


def test_host_planners_clean_under_asan_ubsan(tmp_path):
    exe = tmp_path / "host_sanitize"
    subprocess.run(["g++", "-std=c++17", *SAN, "-I" + CSRC, "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "native", "host_sanitize.cpp"), os.path.join(CSRC, "filter_bank.cpp"),
                    os.path.join(CSRC, "call_plan.cpp"), os.path.join(CSRC, "umma_plan.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1"))
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout + r.stderr


def test_oracle_restatement_clean_under_asan_ubsan(tmp_path):
    """the checker itself: int16 and float entries, capacity-bound calls, three channel counts"""
    src = tmp_path / "drive.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include "speex_oracle.h"
int main(void) {
  static const unsigned cfg[][4] = {{1, 24000, 48000, 5}, {2, 44100, 48000, 7}, {2, 96000, 44100, 10}, {3, 48000, 44100, 4}, {1, 8000, 96000, 2}};
  unsigned long long sum = 0;
  for (unsigned c = 0; c < sizeof cfg / sizeof cfg[0]; ++c) {
    int err = 0;
    orc_resampler *r = orc_create(cfg[c][0], cfg[c][1], cfg[c][2], (int)cfg[c][3], &err);
    if (!r) return 1;
    for (unsigned k = 0; k < 12; ++k) {
      unsigned n = (k * 331u) % 700u, cap = (k * 977u) % 1500u, ch = cfg[c][0];
      short *in = calloc((size_t)n * ch + 1, 2), *out = calloc((size_t)cap * ch + 1, 2);
      float *fin = calloc((size_t)n * ch + 1, 4), *fout = calloc((size_t)cap * ch + 1, 4);
      for (unsigned i = 0; i < n * ch; ++i) { in[i] = (short)((i * 2654435761u) >> 17); fin[i] = in[i] * 0.37f; }
      uint32_t a = n, b = cap;
      if (k & 1) orc_process_interleaved_int16(r, in, &a, out, &b); else orc_process_interleaved_float(r, fin, &a, fout, &b);
      sum += a + b;
      free(in); free(out); free(fin); free(fout);
    }
    orc_destroy(r);
  }
  printf("ok %llu\n", sum);
  return 0;
}
''')
    exe = tmp_path / "drive"
    subprocess.run(["gcc", "-std=c99", *SAN, "-ffp-contract=off", "-I" + os.path.join(ROOT, "oracle"), str(src),
                    os.path.join(ROOT, "oracle", "speex_oracle.c"), "-lm", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout + r.stderr
