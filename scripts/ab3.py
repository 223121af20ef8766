import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, helpers as h
pt = h.pt
sc = h.build_cornell(512, 512); sc.commit()
r = pt.Renderer(sc, 512, 512)
for k in range(3): r.render(0, 64, clear=True)
t=time.perf_counter(); r.render(0, 64, clear=True); wall=(time.perf_counter()-t)*1e3
s = r.stats(); print(os.environ.get('PTB200_LIB','cur')[-12:], 'plain: ms_total %.2f wall %.2f iters %d launches %d rays %d+%d' % (s.ms_total, wall, s.iterations, s.kernel_launches, s.closest_rays, s.shadow_rays))
r.set_flags(pt.FLAG_TIME_KERNELS); r.render(0, 64, clear=True); s = r.stats()
print('   timed: total %.2f gen %.2f trace %.2f shade %.2f shadow %.2f' % (s.ms_total, s.ms_generate, s.ms_trace_closest, s.ms_shade, s.ms_trace_shadow))
