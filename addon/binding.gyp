{
  # node-gyp build of the Node-API addon over libfheb200.so (run on a machine with Node >= 12.17 and the CUDA runtime):
  #   python node-fhe-accelerate_b200/build.py && cd addon && npx node-gyp rebuild
  # node-gyp puts Node's own <node_api.h> on the include path; addon/stub/ is only for the repository's syntax check.
  "targets": [{
    "target_name": "node-fhe-accelerate",
    "sources": ["fheb_addon.cc"],
    "include_dirs": ["../include"],
    "defines": ["NAPI_VERSION=6"],
    "cflags_cc": ["-std=c++17"],
    "libraries": ["-L<(module_root_dir)/../node-fhe-accelerate_b200", "-lfheb200",
                  "-Wl,-rpath,'$$ORIGIN/../../../node-fhe-accelerate_b200'"]
  }]
}
