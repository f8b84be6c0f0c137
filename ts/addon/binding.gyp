{
  "targets": [{
    "target_name": "rt_b200",
    "sources": ["rt_napi.c"],
    "include_dirs": ["../../include"],
    "libraries": ["-L<(module_root_dir)/../../mcp_raytracer_b200/csrc", "-lmcprt_b200", "-Wl,-rpath,<(module_root_dir)/../../mcp_raytracer_b200/csrc"]
  }]
}
