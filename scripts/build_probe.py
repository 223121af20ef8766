import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, helpers as h, facade_dump
pt = h.pt
for name, mk in [('py-mesh708', lambda: h.build_mesh_scene(708, 64, 64)), ('facade-mesh708', lambda: facade_dump.build(pt.HOST_LIB_PATH, 6, 64, 64, 0))]:
    for rep in range(2):
        sc = mk(); d = sc.desc()
        print(name, rep, 'tris', d.n_tris, 'has n/uv/tex', bool(d.tri_n), bool(d.tri_uv), bool(d.tri_tex), flush=True)
        t = time.time(); sc.commit(); print('  commit wall %.1f ms' % ((time.time()-t)*1e3), 'build_ms', sc.bvh_info().build_ms, 'nodes8', sc.bvh_info().n_bvh8_nodes, flush=True)
