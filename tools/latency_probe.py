"""Latency floor of a k_wavefront pass: a zoomed Cornell camera (every pixel sees the inside of the box) on an image so small
that every SM holds one block with `bs` paths -- bs / 32 tasks per pass and SM.  Needs the probe build for the phase split
(SRT_LIB=.../libsrt_prof.so); prints render ms, passes, period and, per kind of first task, the cycles per phase."""
import sys, pathlib
import numpy as np
ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "cuda-spectral-ray-tracer_b200"))
import srt_b200 as S
sc = S.Scene(0)
names = ["regen", "lambert", "metal", "dielectric"]
def run(bs, blocks_per_sm, spp=64):
    n = 148 * blocks_per_sm * bs
    w = 148; h = n // w
    cam = S.CameraBuilder().setVfov(15.0).setLookfrom(278, 278, -800).setLookat(278, 278, 0).setVup(0, 1, 0).getCamera(w, h)
    fb = S.FrameBuffer(w, h)
    rm = S.RenderManager(sc, cam, fb); rm.init_renderer(10, spp)
    rm.set_option(S.OPT_ROUNDS, 1); rm.set_option(S.OPT_PASS_LOG, 1); rm.set_option(S.OPT_BLOCK_SLOTS, bs)
    rm.set_option(S.OPT_TILE_W, 32); rm.set_option(S.OPT_TILE_H, 1)
    rm.init_device_params(0, 0)
    for rep in range(2):
        rm.restart()
        while rm.step(): pass
    st = rm.stats()
    log = rm.pass_log()
    L = log[0]; C = log[4]
    npass = int((L[:, 0] != 0).sum())
    t = L[:npass, 0].astype(np.int64); dt = (np.diff(t) & 0xFFFFFFFF) * 1e-3
    live = (L[:npass, 1] + L[:npass, 2] + (L[:npass, 3] & 0xFFFF) + (L[:npass, 3] >> 16)).astype(np.int64)
    full = live[:-1] >= bs * 3 // 4
    print("bs %4d x %d blocks/SM (%dx%d px, %d spp): render %.3f ms, wavefront launches %d, block 0: %d passes, period while >= 3/4 full %.2f us (%d passes), overall %.2f us; rays/sample %.2f" % (
        bs, blocks_per_sm, w, h, spp, st["render_ms"], st["wavefront_launches"], npass, dt[full].mean() if full.any() else 0, full.sum(), dt.mean(), st["rays"] / max(1, st["samples"])))
    kind = (C[:npass, 3] >> 28).astype(np.int64)
    ph = np.stack([C[:npass, 0], C[:npass, 1], C[:npass, 2], C[:npass, 3] & 0x0FFFFFFF], 1).astype(np.float64)
    ok = (ph.sum(1) > 0)[:-1] & full
    for k in range(4):
        mk = ok & (kind[:-1] == k)
        if mk.sum() < 3: continue
        p = ph[:-1][mk].mean(0)
        print("      first task %-10s (%4d): fetch %5.0f  regen/scatter %5.0f  closest hit %5.0f  store+push %5.0f  = %6.0f of %6.0f cycles per pass" % (
            names[k], mk.sum(), p[0], p[1], p[2], p[3], p.sum(), dt[mk].mean() * 1965))
    F = log[7][:16]
    fine = ["pre-test (flat_candidates)", "exact tests", "extend: box vote", "extend: closest_hit", "extend: sample ends (spectrum + film)", "extend: hit bookkeeping",
            "scatter: normal + unit(d) (incl. wait for the state loads)", "scatter: material branch (RNG / Sellmeier)", "scatter: mul_spectrum"]
    for k, nm in enumerate(fine):
        cyc = float(F[k, 0]) + float(F[k, 1]) * 4294967296.0
        if F[k, 2]:
            print("      fine %-60s %8.0f cycles x %6d visits" % (nm, cyc / F[k, 2], F[k, 2]))
import os
for bs, bps in [(64, 1), (256, 1), (256, 4), (1024, 4)]:
    run(bs, bps)
