{
  # node-gyp build of the N-API shim; libspeexb200.so is built separately by
  # `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a) and found at run time
  # through the rpath below.
  "targets": [
    {
      "target_name": "speexb200",
      "sources": ["src/addon.c"],
      "include_dirs": ["../../include"],
      "defines": ["NAPI_VERSION=6"],
      "libraries": [
        "-L<(module_root_dir)/../../node_speex_resampler_b200",
        "-lspeexb200",
        "-Wl,-rpath,<(module_root_dir)/../../node_speex_resampler_b200"
      ],
      "cflags": ["-O2", "-Wall", "-Wextra"]
    }
  ]
}
