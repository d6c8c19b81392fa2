#!/usr/bin/env python
"""bench.py -- 8K stereo equirect frames/s of the Surround360 render path on B200.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 17-camera
rig, 2048x2048 synthetic frames -> 8K stereo (left over right) equirect, batch_process_video's
"8k" preset (eqr 8400x4096 -> 8192x8192, sharpening 0.25), --enable_top --enable_bottom,
pixflow_low, prev_frame_data_dir NONE. A step = one batch of F independent frames (F = frames in flight per GPU, up to 20, bounded by
free HBM); `value` is frames/s. With N GPUs each rank renders its own batches (weak scaling); the
only collective is one NCCL broadcast of the serialised rig at start-up.

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code (oracle/_ref) on the host cores, same config
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import zlib

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# frames in flight run on separate streams; the default of 8 hardware work queues would serialise them
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np

METRIC = "8k_stereo_equirect_frames_per_sec"
ALGO_BYTES_PER_LEVELPX_PER_SWEEP = 48.0  # "2 sweeps 96 B / level-px"
OUT_MP_PER_FRAME = 8192 * 8192 / 1e6


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full` capture
    (profiles/r2_dominant_kernel_traffic.json, written by tools/ncu_table.py --traffic from the raw CSV next to it); None without it."""
    p = os.path.join(ROOT, "profiles", "r2_dominant_kernel_traffic.json")
    try:
        j = json.load(open(p))
        return float(j["dram_bytes_per_launch"]), j.get("note", "")
    except Exception:
        return None, "no ncu capture committed for this tree"


def preset_kwargs(name, cubemap=False):
    from surround360_b200 import api
    kw = dict(api.PRESETS[name])
    kw.update(sharpening=0.25, enable_top=1, enable_bottom=1)
    if cubemap:
        kw.update(want_cubemap=1, cubemap_width=1536, cubemap_height=1536)
    return kw


def make_frames(rig_path, n_sets):
    """n_sets frame sets: the procedural world rendered once, later sets are the same views
    shifted horizontally so that consecutive steps never see identical inputs."""
    from surround360_b200 import synth
    base = synth.render_rig_frames(rig_path)
    sets = [base]
    for f in range(1, n_sets):
        sets.append([np.ascontiguousarray(np.roll(im, 7 * f, axis=1)) for im in base])
    return sets


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------
# The CPU arm: the reference's own implementation of the path on the host cores.
#   kind "reference": oracle/_ref/libref.so = TestRenderStereoPanorama.cpp, PixFlow.h, NovelView.cpp, CvUtil.cpp, Filter.h,
#                     ImageWarper.cpp, PoleRemoval.cpp, Camera.cpp ... compiled UNCHANGED from /root/reference (oracle/ref_build.sh;
#                     the absent OpenCV primitives come from the cv2-pinned oracle/ocv.h);
#   kind "port":      the oracle restatement (oracle/liboracle.so), used only when oracle/_ref was not built.
# Same preset and flags as the CUDA arm (never a smaller preset). A step = one frame; the reference threads a frame the way
# TestRenderStereoPanorama does (one thread per camera pair, 4 pole threads, 2 sharpen threads), which leaves most of a big
# host idle, so P frames are rendered concurrently (one process each) to use all the host threads it can.
# ---------------------------------------------------------------------------------------------------------------------
def cpu_arm_kind():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref_bind
    return "reference" if ref_bind.available() else "port"


def cpu_arm_worker(args):
    """One worker process: renders frames on request. Protocol on stdin/stdout: 'READY' -> ('GO n' -> 'DONE seconds maxrss_kb')*."""
    import resource
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    kw = preset_kwargs(args.preset, args.cubemap)
    frames = list(np.load(args.frames_npy, mmap_mode="r"))
    if args.worker_kind == "reference":
        import ref_bind as rb
        p = rb.make_params(**kw)
        render = lambda: rb.Frame(args.rig, frames, p).close()
    else:
        import oracle_bind as ob
        rig = ob.Rig.load(args.rig)
        p = ob.make_params(threads=1, **kw)
        render = lambda: rig.render_frame(frames, p)
    print("READY", flush=True)
    for line in sys.stdin:
        f = line.split()
        if not f or f[0] != "GO":
            break
        t0 = time.perf_counter()
        for _ in range(int(f[1])):
            render()
        print("DONE", time.perf_counter() - t0, resource.getrusage(resource.RUSAGE_SELF).ru_maxrss, flush=True)


def run_cpu_arm(preset, cubemap, steps, warmup, rig_path, frames, max_procs=None):
    """-> dict(frames_per_s, seconds, procs, kind, maxrss_gb). `steps` timed frames after `warmup` untimed ones."""
    kind = cpu_arm_kind()
    cores = os.cpu_count() or 1
    try:
        import psutil
        mem_gb = psutil.virtual_memory().available / 2**30
    except Exception:
        mem_gb = 64.0
    per_frame_gb = {"8k": 24.0, "6k": 14.0, "4k": 5.0, "3k": 5.0}[preset]      # resident set of one reference frame, with margin (8k measured: see DESIGN.md)
    procs = max(1, min(steps, cores // 10, int(mem_gb * 0.6 // per_frame_gb), max_procs or 10**6))
    npy = f"/dev/shm/s360_cpu_arm_frames_{os.getpid()}.npy" if os.path.isdir("/dev/shm") else f"/tmp/s360_cpu_arm_frames_{os.getpid()}.npy"
    np.save(npy, np.stack(frames))
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-arm-worker", "--preset", preset, "--worker-kind", kind, "--rig", rig_path,
           "--frames-npy", npy] + (["--cubemap"] if cubemap else [])
    workers = [subprocess.Popen(cmd, stdin=subprocess.PIPE, stdout=subprocess.PIPE, text=True, bufsize=1) for _ in range(procs)]
    try:
        for w in workers:
            if w.stdout.readline().strip() != "READY":
                raise RuntimeError("cpu arm worker failed to start")

        def round_(total):
            per = [total // procs + (1 if i < total % procs else 0) for i in range(procs)]
            t0 = time.perf_counter()
            for w, n in zip(workers, per):
                if n:
                    w.stdin.write(f"GO {n}\n"); w.stdin.flush()
            rss = 0
            for w, n in zip(workers, per):
                if n:
                    f = w.stdout.readline().split()
                    if not f or f[0] != "DONE":
                        raise RuntimeError("cpu arm worker died")
                    rss = max(rss, int(f[2]))
            return time.perf_counter() - t0, rss

        if warmup:
            round_(warmup)
        dt, rss = round_(steps)
    finally:
        for w in workers:
            try:
                w.stdin.close()
            except Exception:
                pass
        for w in workers:
            try:
                w.wait(timeout=30)
            except Exception:
                w.kill()
        try:
            os.remove(npy)
        except OSError:
            pass
    return dict(frames_per_s=steps / dt, seconds=dt, procs=procs, kind=kind, cores=cores, maxrss_gb=rss / 2**20)


def cpu_arm_sample_text(r, preset, steps):
    src = ("the reference's own sources compiled unchanged (oracle/_ref, OpenCV primitives from the cv2-pinned oracle/ocv.h)" if r["kind"] == "reference"
           else "oracle port of the reference (oracle/_ref not built)")
    return (f"{steps} frame(s) of the '{preset}' preset in {r['seconds']:.1f} s wall: {r['procs']} frame(s) at a time, each threaded like "
            f"TestRenderStereoPanorama (one thread per camera pair, 4 pole threads, 2 sharpen threads); {src}")


def run_reference(args, rank, world):
    """`--impl reference`: the CPU arm on the same config, metric and unit as the CUDA arm. Rank 0 only."""
    if rank != 0:
        return
    from surround360_b200 import rig_gen
    rig_path = "/tmp/s360_bench_rig.json"
    rig_gen.write_rig(rig_path)
    frames = make_frames(rig_path, 1)[0]
    kw = preset_kwargs(args.preset, args.cubemap)
    r = run_cpu_arm(args.preset, args.cubemap, args.steps, args.warmup, rig_path, frames)
    value = r["frames_per_s"]
    out_mp = kw["final_eqr_width"] * kw["final_eqr_height"] / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * r["seconds"] / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/f32", "data": "synthetic",
        "config": workload_config(args.preset, args.cubemap, args.inflight, max(world, 1)),
        "mp_per_s": value * out_mp,
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": r["cores"], "kind": r["kind"],
                         "sample": "each step = 1 frame of the step's batch; " + cpu_arm_sample_text(r, args.preset, args.steps),
                         "concurrent_frames": r["procs"], "max_rss_gb_per_frame": r["maxrss_gb"]},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(preset, cubemap, frames_per_step, world):
    """The `config` block: identical in both arms (the reference arm times a bounded sample of this workload)."""
    return {"workload": f"17-cam rig, 2048x2048 synthetic frames -> 8K stereo equirect ({preset} preset, top+bottom, pixflow_low"
                        + (", +cubemap 1536" if cubemap else "") + ")",
            "frames_per_step": frames_per_step, "frames_in_flight_per_gpu": frames_per_step,
            "l2": "inputs (214 MB/frame) exceed the 126 MB L2; two alternating input sets",
            "parallelism": f"frames sharded one-per-GPU x{world}, single NCCL broadcast of the rig"}


def side_pass(api, torch, dev, local_rank, blob_host, frame_sets, preset, cubemap, F, steps, temporal=False):
    """A short device-resident throughput pass of another configuration (frames in flight = F). temporal: video mode -- every context is
    a chain whose frame i regularises against its frame i-1 (--prev_frame_data_dir, kept on the device: s360_ctx_set_temporal)."""
    params = api.default_params(**preset_kwargs(preset, cubemap))
    ctxs = []
    for _ in range(F):
        c = api.Context(local_rank)
        c.load_rig_blob(blob_host)
        if temporal:
            c.set_temporal(True)
        ctxs.append(c)
    eq_w, eq_h, cw, ch = ctxs[0].output_size(params)
    img_h, img_w = frame_sets[0][0].shape[:2]
    dev_in = [[torch.from_numpy(im).to(dev) for im in fs] for fs in frame_sets]
    outs = [torch.empty((eq_h, eq_w, 3), dtype=torch.uint8, device=dev) for _ in range(F)]
    cubes = [torch.empty((max(ch, 1), max(cw, 1), 3), dtype=torch.uint8, device=dev) for _ in range(F)]

    def enqueue(s):
        ctxs[s % F].render_frame_dev(params, [t.data_ptr() for t in dev_in[s % len(dev_in)]], img_w, img_h, outs[s % F].data_ptr(),
                                     cubes[s % F].data_ptr() if cw else None)
    for s in range(3 * F):
        enqueue(s)
    for c in ctxs:
        c.synchronize()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(steps * F):
        enqueue(s)
    for c in ctxs:
        c.synchronize()
    dt = time.perf_counter() - t0
    for c in ctxs:
        c.close()
    torch.cuda.empty_cache()
    fps = steps * F / dt
    r = {"preset": preset, "cubemap": bool(cubemap), "frames_in_flight": F, "steps": steps, "frames_per_s": fps,
         "stereo_mp_per_s": fps * eq_w * eq_h / 1e6, "timing": "host clock around enqueue + synchronize of steps x F frames, inputs resident in HBM"}
    if preset == "6k":
        r["mono_left_eye_mp_per_s"] = fps * eq_w * (eq_h // 2) / 1e6   # BASELINE's "6K mono" = the left eye (the reference always renders both)
    if temporal:
        r["video_mode"] = "every context is one chain of consecutive frames (temporal regularisation against the previous frame, on the device)"
        r["chains_in_flight"] = F
    return r


ISP_CONFIG = {"CameraIsp": {   # a 16-bit GBRG sensor in the schema of the reference's res/config/isp/*.json (values written here)
    "bitsPerPixel": 16, "blackLevel": [1310.0, 1290.0, 1275.0],
    "vignetteRollOffH": [[1.12, 1.10, 1.08], [1.02, 1.01, 1.0], [1.0, 1.0, 1.0], [1.01, 1.02, 1.0], [1.09, 1.11, 1.13]],
    "vignetteRollOffV": [[1.15, 1.12, 1.1], [1.0, 1.0, 1.0], [1.0, 1.0, 1.0], [1.07, 1.09, 1.11]],
    "whiteBalanceGain": [1.17, 1.0, 1.58], "stuckPixelThreshold": 5, "stuckPixelDarknessThreshold": 0.11, "stuckPixelRadius": 0,
    "ccm": [[1.04, -0.06, 0.02], [0.15, 1.12, -0.27], [-0.14, -0.09, 1.23]], "sharpening": [0.5, 0.45, 0.55], "saturation": 1.2, "contrast": 1.0,
    "lowKeyBoost": [-0.2, -0.15, -0.2], "highKeyBoost": [0.2, 0.2, 0.15], "gamma": [0.4545, 0.46, 0.45], "bayerPattern": "GBRG"}}

ISP_CPU_SNIPPET = """
import sys, time, json, zlib
import numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import ref_bind as rb
from surround360_b200 import synth
raw = synth.bayer_frame(2048, 2048, np.uint16, seed=5)
cfg = %r
rb.isp_process(cfg, raw[:256, :256], 2, 1, 16)
t0 = time.perf_counter(); n = 0
while n < 3:
    out = rb.isp_process(cfg, raw, 2, 1, 16); n += 1
print(json.dumps({"seconds_per_image": (time.perf_counter() - t0) / n, "crc32": format(zlib.crc32((out >> 8).astype(np.uint8).tobytes()), "08x")}))
"""


def isp_pass(api, torch, dev, local_rank, peak, cpu=True):
    """the soft ISP in front of the renderer. One 'frame' = the rig's 17 raw 2048 x 2048 16-bit Bayer images -> BGR8 renderer
    input (edge-aware demosaic, tone curve, IIR unsharp mask; 16-bit pipeline narrowed to the high byte like the reference's PNG hand-off),
    device resident, timed with CUDA events on the context's stream. CPU sample: the reference's own CameraIsp (oracle/_ref) on one image,
    one thread (the reference runs one camera per thread, Unpacker.cpp:118), in a subprocess."""
    from surround360_b200 import synth
    ctx = api.Context(local_rank)
    cfg = json.dumps(ISP_CONFIG)
    isp = api.CameraIsp(cfg, 16)
    isp.setBitsPerPixel(16); isp.setDemosaicFilter(2); isp.setResize(1); isp.enableToneMap()
    n_cam, w, h = 17, 2048, 2048
    raws = [torch.from_numpy(synth.bayer_frame(h, w, np.uint16, seed=5 + (k % 3)).view(np.int16)).to(dev) for k in range(n_cam)]
    outs = [torch.empty((h, w, 3), dtype=torch.uint8, device=dev) for _ in range(n_cam)]
    st = torch.cuda.ExternalStream(ctx.stream(), device=dev)

    rp, op = [r.data_ptr() for r in raws], [o.data_ptr() for o in outs]

    def frame():   # the 17 cameras through shared launches (s360_isp_process_frames_dev)
        isp.frameImagesDev(ctx, rp, w, h, 16, op)
    for _ in range(3):
        frame()
    ctx.synchronize()
    K = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(K):
        frame()
    e1.record(st)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / K
    crc = format(zlib.crc32(outs[0].cpu().numpy().tobytes()), "08x")
    # the same frames as the capture software stores them (.bin footage: packed 12-bit samples, Unpacker.cpp:141-190): widened inside the
    # ISP's first kernel; the 12-bit values are the 16-bit frames' high 12 bits, so the output differs from the pass above
    packs = [torch.from_numpy(synth.pack12(synth.bayer_frame(h, w, np.uint16, seed=5 + (k % 3)) >> 4)).to(dev) for k in range(n_cam)]
    pp = [p.data_ptr() for p in packs]
    for _ in range(3):
        isp.footageImagesDev(ctx, pp, 12, w, h, op, high_byte=True)
    ctx.synchronize()
    e0.record(st)
    for _ in range(K):
        isp.footageImagesDev(ctx, pp, 12, w, h, op, high_byte=True)
    e1.record(st)
    ctx.synchronize()
    ms_footage = e0.elapsed_time(e1) / K
    frame()   # the 16-bit pass again: `outs` is what the per-kernel table and the CPU check below refer to
    ctx.synchronize()
    # per-kernel device time of one frame (events around every launch)
    ctx.profile_enable(2)
    frame()
    ctx.synchronize()
    kern = []
    for r in ctx.profile_kernels():
        if r["kernel"].startswith("k_isp"):
            gbs = r["algorithmic_bytes"] / 1e9 / (r["ms"] / 1e3) if r["ms"] > 0 else 0.0
            kern.append({"kernel": r["kernel"], "launches_per_frame": r["launches"], "ms_per_frame_alone": r["ms"], "algorithmic_gb_per_frame": r["algorithmic_bytes"] / 1e9,
                         "achieved_gbs": gbs, "frac": gbs / peak})
    ctx.profile_enable(0)
    isp.close(); ctx.close()
    res = {"workload": "17 raw 2048x2048 uint16 GBRG frames -> BGR8 (edge-aware demosaic, CCM + tone curve, IIR unsharp mask), inputs and outputs resident in HBM",
           "ms_per_rig_frame": ms, "raw_mp_per_s": n_cam * w * h / 1e6 / (ms / 1e3), "rig_frames_per_s": 1000.0 / ms, "launches_per_rig_frame": 10,
           "output_crc32_cam0": crc, "footage12_ms_per_rig_frame": ms_footage, "footage12_raw_mp_per_s": n_cam * w * h / 1e6 / (ms_footage / 1e3),
           "kernels": sorted(kern, key=lambda k: -k["ms_per_frame_alone"])}
    if cpu and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_isp.so")):
        r = subprocess.run([sys.executable, "-c", ISP_CPU_SNIPPET % (os.path.join(ROOT, "tests"), ROOT, cfg)], capture_output=True, text=True, timeout=600)
        if r.returncode == 0:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            res["cpu_baseline"] = {"value": w * h / 1e6 / j["seconds_per_image"], "unit": "raw MP/s", "cores": 1, "kind": "reference",
                                   "sample": "3 x one 2048x2048 image through the reference's CameraIsp (oracle/_ref/libref_isp.so), one thread"}
            res["matches_cpu_reference"] = j["crc32"] == crc   # camera 0 is the image of the CPU sample (seed 5)
    return res


def nv12_pass(api, torch, dev, local_rank, peak, eq_w, eq_h):
    """the finished stereo equirect (BGR8 in HBM) -> the NV12 surface a hardware encoder registers, on the device
    (s360_frame_to_nv12_dev); CUDA events on the context's stream; algorithmic bytes 3 in + 1.5 out per pixel."""
    ctx = api.Context(local_rank)
    g = torch.Generator(device="cpu"); g.manual_seed(11)
    frame = torch.randint(0, 256, (eq_h, eq_w, 3), dtype=torch.uint8, generator=g).to(dev)
    y = torch.empty((eq_h, eq_w), dtype=torch.uint8, device=dev)
    uv = torch.empty((eq_h // 2, eq_w), dtype=torch.uint8, device=dev)
    st = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    for _ in range(3):
        ctx.frame_to_nv12_dev(frame.data_ptr(), eq_w, eq_h, y.data_ptr(), uv.data_ptr(), matrix=1)
    ctx.synchronize()
    K = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(K):
        ctx.frame_to_nv12_dev(frame.data_ptr(), eq_w, eq_h, y.data_ptr(), uv.data_ptr(), matrix=1)
    e1.record(st)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / K
    gbs = eq_w * eq_h * 4.5 / 1e9 / (ms / 1e3)
    crc = format(zlib.crc32(y.cpu().numpy().tobytes() + uv.cpu().numpy().tobytes()), "08x")
    ctx.close()
    return {"workload": f"{eq_w}x{eq_h} BGR8 stereo equirect -> NV12 (BT.709 limited range), resident in HBM; the frame is 201 MB, larger than the L2",
            "ms_per_frame": ms, "achieved_gbs": gbs, "frac": gbs / peak, "nv12_crc32": crc,
            "note": "the image has no encoder library: the hand-off ends at the surface NVENC would register (INTEGRATION.md)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--preset", default="8k", choices=["3k", "4k", "6k", "8k"])
    ap.add_argument("--cubemap", action="store_true", help="config 5: also produce the 1536^2 cubemap strip")
    ap.add_argument("--inflight", type=int, default=23, help="frames in flight per GPU (contexts/streams); 23 contexts of 6.6 GiB fit in 180 GB at the 8k preset; reduced further if fewer fit")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--temporal", action="store_true", help="video mode: every context is a chain (frame i regularises against frame i-1 of its context)")
    ap.add_argument("--no-extras", action="store_true", help="skip the short 6k / cubemap passes appended to the headline line")
    ap.add_argument("--latency-mode", action="store_true", help="run every context in latency mode (s360_ctx_set_latency_mode); the default is the throughput mode")
    ap.add_argument("--isp-only", action="store_true", help="only the soft-ISP pass of the extras (SURVEY 8 f2): prints its JSON object")
    ap.add_argument("--cpu-arm-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--worker-kind", default="reference", help=argparse.SUPPRESS)
    ap.add_argument("--rig", default="", help=argparse.SUPPRESS)
    ap.add_argument("--frames-npy", default="", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_arm_worker:
        cpu_arm_worker(args)
        return
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from surround360_b200 import api, rig_gen
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if args.isp_only:
        print(json.dumps(isp_pass(api, torch, dev, local_rank, peaks()[0], cpu=not args.no_cpu_baseline)), flush=True)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    F = max(1, args.inflight)
    ctxs = [api.Context(local_rank)]
    if args.temporal:
        ctxs[0].set_temporal(True)
    # rig: parsed on rank 0, one NCCL broadcast of the serialised tables, every rank deserialises
    rig_path = f"/tmp/s360_bench_rig_{rank}.json"
    rig_gen.write_rig(rig_path)
    if rank == 0:
        ctxs[0].load_rig(rig_path)
        blob = torch.from_numpy(ctxs[0].rig_blob()).to(dev)
        size = torch.tensor([blob.numel()], device=dev, dtype=torch.int64)
    else:
        size = torch.zeros(1, device=dev, dtype=torch.int64)
    if world > 1:
        dist.broadcast(size, 0)
        if rank != 0:
            blob = torch.zeros(int(size.item()), device=dev, dtype=torch.uint8)
        dist.broadcast(blob, 0)
    blob_host = blob.cpu().numpy()
    ctxs[0].load_rig_blob(blob_host)

    params = api.default_params(**preset_kwargs(args.preset, args.cubemap))
    eq_w, eq_h, cw, ch = ctxs[0].output_size(params)
    n_sets = 2
    frame_sets = make_frames(rig_path, n_sets)
    # frames in flight are bounded by HBM: build the first context's plan, measure what it took and keep
    # 12 % of the device free
    free0, total_mem = torch.cuda.mem_get_info(dev)
    ctxs[0].renderStereoPanorama(params, frame_sets[0])
    free1, _ = torch.cuda.mem_get_info(dev)
    per_ctx = max(free0 - free1, 1)
    fit = int((free1 - 0.12 * total_mem) // per_ctx) + 1
    if fit < F:
        print(f"[bench] {F} frames in flight do not fit ({per_ctx / 2**30:.1f} GiB per context): using {max(fit, 1)}", file=sys.stderr)
        F = max(fit, 1)
    for _ in range(F - 1):
        c = api.Context(local_rank)
        c.load_rig_blob(blob_host)
        ctxs.append(c)
    if args.temporal:
        for c in ctxs[1:]:
            c.set_temporal(True)
    if args.latency_mode:
        for c in ctxs:
            c.set_latency_mode(True)
    img_h, img_w = frame_sets[0][0].shape[:2]
    n_cam = len(frame_sets[0])
    h2d_bytes = n_cam * img_w * img_h * 3
    d2h_bytes = eq_w * eq_h * 3 + cw * ch * 3

    # device-resident inputs/outputs for `value`
    dev_in = [[torch.from_numpy(im).to(dev) for im in fs] for fs in frame_sets]
    dev_out = [torch.empty((eq_h, eq_w, 3), dtype=torch.uint8, device=dev) for _ in range(F)]
    dev_cube = [torch.empty((max(ch, 1), max(cw, 1), 3), dtype=torch.uint8, device=dev) for _ in range(F)]
    ext = [torch.cuda.ExternalStream(c.stream(), device=dev) for c in ctxs]

    def enqueue(step):
        c = ctxs[step % F]
        ins = dev_in[step % n_sets]
        c.render_frame_dev(params, [t.data_ptr() for t in ins], img_w, img_h, dev_out[step % F].data_ptr(),
                           dev_cube[step % F].data_ptr() if cw else None)

    def sync_all():
        for c in ctxs:
            c.synchronize()
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    # A step = one batch of F frames (one per context in flight), so that the measured throughput does not depend
    # on how many steps the caller asks for. Warm-up batches also build every context's plan and frame graph.
    for s in range(max(args.warmup, 1) * F):
        enqueue(s)
    sync_all()

    def timed_pass(n_steps):
        """n_steps frames, F in flight; device time between the first enqueue and the last completion."""
        barrier(); sync_all()
        launches_before = ctxs[0].launch_count()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        start.record(cur)
        for e in ext:
            e.wait_event(start)
        t_enq = time.perf_counter()
        for s in range(n_steps):
            enqueue(s)
        t_enq = time.perf_counter() - t_enq
        for e in ext:
            ev = torch.cuda.Event()
            ev.record(e)
            cur.wait_event(ev)
        end.record(cur)
        end.synchronize()
        sync_all(); barrier()
        return start.elapsed_time(end), ctxs[0].launch_count() - launches_before, t_enq

    # pass A: the measured throughput (whole-frame CUDA-graph replay, no instrumentation)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches, t_enq = timed_pass(args.steps * F)
    clocks = sampler.stop()
    # pass B: the same region with CUDA events around every sweep launch (plain launch path) for the roofline block
    for c in ctxs:
        c.profile_enable(True)
    prof_steps = min(args.steps, 2)
    ms_prof, _, _ = timed_pass(prof_steps * F)
    prof = dict(ms=0.0, launches=0, level_pixels=0.0, steps=0.0, fallbacks=0)
    for c in ctxs:
        pr = c.profile_sweep()
        for k in prof:
            prof[k] += pr[k]
        c.profile_enable(False)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * args.steps * F / (ms_max / 1000.0)

    # stage breakdown of one frame alone on the device (host API fills the stage timers)
    pinned_in = [[torch.from_numpy(im).pin_memory() for im in fs] for fs in frame_sets]
    pinned_out = [torch.empty((eq_h, eq_w, 3), dtype=torch.uint8).pin_memory() for _ in range(F)]
    pinned_cube = [torch.empty((max(ch, 1), max(cw, 1), 3), dtype=torch.uint8).pin_memory() for _ in range(F)]
    ctxs[0].set_graphs(False)
    ctxs[0].renderStereoPanorama(params, [t.numpy() for t in pinned_in[0]], pinned_out[0].numpy(), pinned_cube[0].numpy() if cw else None)
    stage_ms = ctxs[0].stage_ms()
    ctxs[0].set_graphs(True)
    # the same frame alone in the other sweep mode (s360_ctx_set_latency_mode: cp.async input / footprint rings -- shorter alone, slower
    # with many frames in flight, bit-identical): plan rebuilt, one warm frame, one stage-timed frame, CRC against the default mode's frame
    import zlib
    crc_mode_a = zlib.crc32(pinned_out[0].numpy().tobytes())
    other = not args.latency_mode
    ctxs[0].set_latency_mode(other)
    ctxs[0].renderStereoPanorama(params, [t.numpy() for t in pinned_in[0]], pinned_out[0].numpy(), pinned_cube[0].numpy() if cw else None)
    ctxs[0].set_graphs(False)
    ctxs[0].renderStereoPanorama(params, [t.numpy() for t in pinned_in[0]], pinned_out[0].numpy(), pinned_cube[0].numpy() if cw else None)
    stage_ms_other = ctxs[0].stage_ms()
    ctxs[0].set_graphs(True)
    if zlib.crc32(pinned_out[0].numpy().tobytes()) != crc_mode_a:
        raise SystemExit("bench: latency mode and throughput mode rendered different frames")
    ctxs[0].set_latency_mode(args.latency_mode)
    for _ in range(2):   # rebuild the plan and re-capture the frame graph of the mode the rest of the run uses
        ctxs[0].render_frame_dev(params, [t.data_ptr() for t in dev_in[0]], img_w, img_h, dev_out[0].data_ptr(), dev_cube[0].data_ptr() if cw else None)
    ctxs[0].synchronize()
    single_frame = {"sweep_mode_of_this_run": "latency" if args.latency_mode else "throughput",
                    "throughput_mode_ms": (stage_ms_other if args.latency_mode else stage_ms)["total"],
                    "latency_mode_ms": (stage_ms if args.latency_mode else stage_ms_other)["total"],
                    "latency_mode_stages": stage_ms if args.latency_mode else stage_ms_other,
                    "note": "one 8K frame alone on the GPU through the host API (stage timers); s360_ctx_set_latency_mode selects the sweep variant, results are bit-identical"}

    # every kernel of one frame rendered ALONE on the GPU with CUDA events around each launch (s360_profile_kernels): ms per frame,
    # algorithmic bytes per frame (compulsory reads + writes) and the fraction of the HBM roofline that gives
    ctxs[0].profile_enable(2)      # level 2: events around every kernel launch
    n_prof_frames = 2
    for i in range(n_prof_frames):
        ctxs[0].render_frame_dev(params, [t.data_ptr() for t in dev_in[i % n_sets]], img_w, img_h, dev_out[0].data_ptr(), dev_cube[0].data_ptr() if cw else None)
    kernel_table = ctxs[0].profile_kernels()
    ctxs[0].profile_enable(False)

    # the graph-replay path that `value` times and the host-pointer path that `e2e` times (and the GPU tests compare with the oracle
    # and the reference build at this size) must give the same frame: CRC32 of frame 0 through both
    import zlib
    ctxs[0].render_frame_dev(params, [t.data_ptr() for t in dev_in[0]], img_w, img_h, dev_out[0].data_ptr(), dev_cube[0].data_ptr() if cw else None)
    ctxs[0].synchronize()
    crc_graph = zlib.crc32(dev_out[0].cpu().numpy().tobytes())
    ctxs[0].renderStereoPanorama(params, [t.numpy() for t in pinned_in[0]], pinned_out[0].numpy(), pinned_cube[0].numpy() if cw else None)
    crc_host = zlib.crc32(pinned_out[0].numpy().tobytes())
    if crc_graph != crc_host:
        raise SystemExit(f"bench: the graph-replay frame (crc {crc_graph:08x}) differs from the host-API frame (crc {crc_host:08x})")

    # e2e: the public host-pointer API, pinned host inputs copied in and the result copied out every step
    e2e = None
    if not args.no_e2e:
        def worker(i, n_steps, offset):
            c = ctxs[i]
            for s in range(n_steps):
                fs = pinned_in[(offset + s * F + i) % n_sets]
                c.renderStereoPanorama(params, [t.numpy() for t in fs], pinned_out[i].numpy(), pinned_cube[i].numpy() if cw else None)
        def run_threads(total):
            per = [total // F + (1 if i < total % F else 0) for i in range(F)]
            th = [threading.Thread(target=worker, args=(i, per[i], 0)) for i in range(F) if per[i]]
            for t_ in th: t_.start()
            for t_ in th: t_.join()
        run_threads(F)  # warm
        barrier(); sync_all()
        t0 = time.perf_counter()
        run_threads(args.steps * F)
        sync_all()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * args.steps * F / float(tt.item()), "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes * F,
               "d2h_bytes_per_step": d2h_bytes * F}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    traffic, traffic_note = ncu_traffic()
    roof = None
    if prof["launches"]:
        per_launch_s = prof["ms"] / 1000.0 / prof["launches"]
        bytes_per_launch = ALGO_BYTES_PER_LEVELPX_PER_SWEEP * prof["level_pixels"] / prof["launches"]
        achieved = bytes_per_launch / per_launch_s / 1e9
        roof = {"bound": "hbm", "kernel": "k_sweep (PixFlow wavefront sweep, forward+backward)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peak_src,
                "avg_launch_ms": prof["ms"] / prof["launches"], "launches_timed": prof["launches"],
                "ns_per_wavefront_step": 1e6 * prof["ms"] / max(prof["steps"], 1.0),
                "sweep_ms_per_frame": prof["ms"] / (prof_steps * F), "sweep_generic_reruns_last_frames": prof["fallbacks"], "instrumented_pass_ms_per_step": ms_prof / prof_steps,
                "note": "dependency-latency-bound recurrence: reported against the HBM roofline for completeness"}

    roofline_all = []
    for r in sorted(kernel_table, key=lambda r: -r["ms"]):
        ms_f = r["ms"] / n_prof_frames
        gb = r["algorithmic_bytes"] / n_prof_frames / 1e9
        ach = gb / (ms_f / 1e3) if ms_f > 0 else 0.0
        roofline_all.append({"kernel": r["kernel"], "launches_per_frame": r["launches"] / n_prof_frames, "ms_per_frame_alone": ms_f,
                             "algorithmic_gb_per_frame": gb, "achieved_gbs": ach, "frac": ach / peak})

    extra = None
    if not args.no_extras and world == 1 and args.preset == "8k" and not args.cubemap:
        # BASELINE configs 2 and 5 at HEAD, a few steps each: release the 8K contexts first (HBM)
        for c in ctxs:
            c.close()
        ctxs.clear()
        torch.cuda.empty_cache()
        extra = {}
        for name, preset, cube in (("config2_6k_mono", "6k", False), ("config5_8k_cubemap", "8k", True)):
            extra[name] = side_pass(api, torch, dev, local_rank, blob_host, frame_sets, preset, cube, min(F, 12), steps=2)
        # video mode (SURVEY 8e option A: contiguous frame ranges, one chain per context): many chains in flight, and one chain alone
        extra["video_8k_chains"] = side_pass(api, torch, dev, local_rank, blob_host, frame_sets, "8k", False, min(F, 12), steps=2, temporal=True)
        extra["video_8k_single_chain"] = side_pass(api, torch, dev, local_rank, blob_host, frame_sets, "8k", False, 1, steps=4, temporal=True)
        extra["isp_raw_to_bgr"] = isp_pass(api, torch, dev, local_rank, peak, cpu=not args.no_cpu_baseline)
        extra["encode_handoff_nv12"] = nv12_pass(api, torch, dev, local_rank, peak, eq_w, eq_h)

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        # bounded sample on the host cores, in worker processes of their own (the GPU arm's process never loads the CPU code):
        # one round of concurrent frames of the same preset
        P = max(1, min((os.cpu_count() or 1) // 16, 8))
        r = run_cpu_arm(args.preset, args.cubemap, P, 0, rig_path, frame_sets[0], max_procs=P)
        cpu_baseline = {"value": r["frames_per_s"], "unit": "frames/s", "cores": r["cores"], "kind": r["kind"],
                        "sample": cpu_arm_sample_text(r, args.preset, P), "concurrent_frames": r["procs"], "max_rss_gb_per_frame": r["maxrss_gb"]}

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/f32", "data": "synthetic",
        "config": dict(workload_config(args.preset, args.cubemap, F, world), **({"video_mode": "temporal chains, one per context"} if args.temporal else {})),
        "mp_per_s": value * eq_w * eq_h / 1e6,
        "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e, "roofline": roof, "roofline_all": roofline_all, "cpu_baseline": cpu_baseline,
        "output_crc32": {"graph_replay_frame0": f"{crc_graph:08x}", "host_api_frame0": f"{crc_host:08x}", "equal": True},
        "extra": extra,
        "stage_ms_single_frame": stage_ms, "single_frame_latency": single_frame, "host_enqueue_ms_per_frame": 1000.0 * t_enq / (args.steps * F),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
