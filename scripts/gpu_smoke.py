"""Dev script: first contact with the GPU — parity + rough timing (not the bench)."""
import sys, time, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np
import helpers as h
pt = h.pt

def check_scene(name, sc, W, H, nf, big=False):
    t = time.time(); sc.commit(); tc = time.time() - t
    bi = sc.bvh_info()
    print(f"[{name}] commit {tc*1e3:.1f} ms build {bi.build_ms:.2f} ms prims {bi.n_prims} nodes8 {bi.n_bvh8_nodes} sah {bi.sah_cost_lbvh:.2f}->{bi.sah_cost_refined:.2f}", flush=True)
    orc = h.OracleScene(sc)
    rays = orc.camera_rays(W, H, 0)
    rays = np.concatenate([rays, orc.bounce_rays(rays, 7)])
    ho, _ = orc.trace(rays)
    hg, cnt = sc.trace(rays, count_fetches=True)
    same = (ho['prim'] == hg['prim']) & (ho['t'].view(np.uint32) == hg['t'].view(np.uint32)) & (ho['u'].view(np.uint32) == hg['u'].view(np.uint32))
    print(f"   trace: {len(rays)} rays, hit frac {(ho['prim']>=0).mean():.3f}, mismatches {(~same).sum()}, nodes/ray {cnt[0]/len(rays):.2f} leaves/ray {cnt[1]/len(rays):.2f}", flush=True)
    r = pt.Renderer(sc, W, H)
    r.render(0, nf)
    g = r.read_accum()
    st = r.stats()
    if not big:
        a, ost = orc.render(W, H, 0, nf)
        d = np.abs(a - g)[..., :3].max(axis=-1)
        print(f"   render: maxdiff {d.max():.3g} pixels>1e-3: {(d>1e-3).sum()}/{W*H} w_ok {(a[...,3]==g[...,3]).all()} mean {a[...,:3].mean():.5f} vs {g[...,:3].mean():.5f}; oracle rays {ost.closest_rays}+{ost.shadow_rays} gpu {st.closest_rays}+{st.shadow_rays}", flush=True)
    print(f"   gpu: {st.ms_total:.2f} ms, {st.samples/st.ms_total/1e3:.1f} Msamples/s, {(st.closest_rays+st.shadow_rays)/st.ms_total/1e3:.1f} Mrays/s, iters {st.iterations} launches {st.kernel_launches}", flush=True)
    return r

check_scene('cornell', h.build_cornell(128, 128), 128, 128, 8)
check_scene('zoo', h.build_material_zoo(128, 128), 128, 128, 8)
check_scene('spheres', h.build_spheres(22, W=128, H=128), 128, 128, 8)
check_scene('soup', h.build_random_prims(1, W=64, H=64), 64, 64, 2)
check_scene('mesh', h.build_mesh_scene(64, 128, 128), 128, 128, 8)
check_scene('meshN', h.build_mesh_scene(64, 128, 128, normals=True, mat_kind=pt.MAT_METAL), 128, 128, 4)
# larger timing runs
r = check_scene('spheres-big', h.build_spheres(22, W=1200, H=800), 1200, 800, 16, big=True)
r = check_scene('cornell-big', h.build_cornell(512, 512), 512, 512, 64, big=True)
r = check_scene('mesh-1M', h.build_mesh_scene(708, 1920, 1080), 1920, 1080, 4, big=True)
