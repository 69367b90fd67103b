import sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
w,h,spp=1920,1080,64
def T(): return time.perf_counter()
for rep in range(4):
    t0=T(); sc=S.Scene(0); t1=T(); fb=S.FrameBuffer(w,h); t2=T(); rm=S.RenderManager(sc, sc.camera(w,h), fb); rm.init_renderer(10,spp); t3=T()
    rm.init_device_params(0,0); t4=T(); rm.render_all(); t5=T(); st=rm.stats(); del rm; t6=T(); del sc; t7=T()
    print("scene %.1f  fb %.1f  rm+init_renderer %.1f  init_device_params %.1f  render_all %.1f (kernel %.1f)  del rm %.1f  del scene %.1f ms" % tuple(1e3*x for x in (t1-t0,t2-t1,t3-t2,t4-t3,t5-t4,st["render_ms"]*1e-3,t6-t5,t7-t6)))
