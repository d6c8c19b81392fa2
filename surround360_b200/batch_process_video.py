#!/usr/bin/env python
"""batch_process_video for the B200 path: the caller of the frame renderer (reference: surround360_render/scripts/batch_process_video.py).

The reference script renders frames one after the other, one TestRenderStereoPanorama process per frame, every frame regularising its
flows against the previous frame's through files (--prev_frame_data_dir, scripts/batch_process_video.py:123-210). Here the frames of
a sequence are rendered in long-lived contexts with the previous frame's state on the device (s360_ctx_set_temporal), several chains at a
time per GPU and one process per GPU:

  python -m surround360_b200.batch_process_video --root_dir R --start_frame 0 --end_frame 299 --quality 8k --rig_json_file rig.json \\
      --flow_alg pixflow_low --enable_top --enable_bottom [--gpus 8] [--chains 12] [--independent] [--dryrun]

Same arguments as the reference script (root_dir layout: rgb/<camera id>/<frame:06>.png in, eqr_frames/eqr_<frame:06>.png and
cube_frames/cube_<frame:06>.png out; quality presets 3k / 4k / 6k / 8k; sharpening 0.25; --cubemap_width / --cubemap_height /
--cubemap_format; --enable_pole_removal with root_dir/pole_masks/<camera id>.png), plus:
  --gpus N        processes (one per GPU; also taken from torchrun's WORLD_SIZE / RANK / LOCAL_RANK)
  --chains C      chains in flight per GPU. The sequence is cut into N x C contiguous ranges (SURVEY 8e option A, dist.chain_ranges); the
                  first frame of a range starts without history, like the first frame of a video, so the output deviates from a
                  single-chain render at those N x C - 1 range heads (--gpus 1 --chains 1 is the reference's single chain)
  --independent   no temporal regularisation at all (every frame as with --prev_frame_data_dir NONE): frame f -> rank f mod N
  --dryrun        print the plan (which rank / chain renders which frames) and exit; needs no GPU
Not supported: --resume (the reference reads the previous frame's flow files; this driver keeps that state on the device and writes no
flow files -- use the TestRenderStereoPanorama tool for file-based chains), --save_debug_images (accepted, ignored with a warning)."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

from . import dist

SHARPENING = 0.25   # scripts/batch_process_video.py:177-199: every preset


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="batch process video frames (B200 path)")
    p.add_argument("--root_dir", required=True, help="path to frame container dir")
    p.add_argument("--surround360_render_dir", default=".", help="accepted for compatibility (no binaries are launched)")
    p.add_argument("--start_frame", required=True, type=int)
    p.add_argument("--end_frame", required=True, type=int)
    p.add_argument("--quality", required=True, help="3k,4k,6k,8k")
    p.add_argument("--cubemap_width", default=0, type=int)
    p.add_argument("--cubemap_height", default=0, type=int)
    p.add_argument("--cubemap_format", default="photo", help="photo,video")
    p.add_argument("--save_debug_images", action="store_true")
    p.add_argument("--enable_top", action="store_true")
    p.add_argument("--enable_bottom", action="store_true")
    p.add_argument("--enable_pole_removal", action="store_true")
    p.add_argument("--resume", action="store_true")
    p.add_argument("--rig_json_file", required=True)
    p.add_argument("--flow_alg", required=True, help="flow algorithm e.g., pixflow_low, pixflow_search_20")
    p.add_argument("--verbose", action="store_true")
    p.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    p.add_argument("--chains", type=int, default=12, help="chains (contexts) in flight per GPU")
    p.add_argument("--independent", action="store_true", help="no temporal regularisation: frames are independent units")
    p.add_argument("--dryrun", action="store_true")
    return p.parse_args(argv)


def quality_params(quality):
    from . import api
    if quality not in api.PRESETS:
        raise SystemExit("Unrecognized quality setting: " + quality)   # scripts/batch_process_video.py:200-202
    return dict(api.PRESETS[quality], sharpening=SHARPENING)


def make_plan(args):
    """-> plan[rank] = list of chains; a chain = dict(frames=[absolute frame numbers in render order], temporal=bool)."""
    if args.enable_pole_removal and not args.enable_bottom:
        raise SystemExit("Cannot use enable_pole_removal if enable_bottom is not used")   # :164-166
    if args.resume:
        raise SystemExit("--resume is not supported by this driver (the previous frame's state lives on the device): "
                         "use TestRenderStereoPanorama --prev_frame_data_dir for chains that continue from files")
    if args.gpus < 1 or args.chains < 1:
        raise SystemExit("--gpus and --chains must be positive")
    frames = list(range(args.start_frame, args.end_frame + 1))
    plan = []
    if args.independent:
        for r in range(args.gpus):
            mine = [frames[i] for i in dist.frames_for_rank(len(frames), r, args.gpus)]
            plan.append([dict(frames=mine[c::args.chains], temporal=False) for c in range(args.chains) if mine[c::args.chains]])
    else:
        ranges = dist.chain_ranges(len(frames), args.gpus, args.chains)
        for r in range(args.gpus):
            plan.append([dict(frames=frames[a:b], temporal=True) for (a, b) in ranges[r] if b > a])
    return plan


def describe(plan):
    lines = []
    for r, chains in enumerate(plan):
        for c, ch in enumerate(chains):
            f = ch["frames"]
            span = f"{f[0]:06d}..{f[-1]:06d}" if ch["temporal"] else ",".join(f"{x:06d}" for x in f[:6]) + (",..." if len(f) > 6 else "")
            lines.append(f"rank {r} chain {c}: {len(f)} frame(s) {span}" + (" (temporal chain, head without history)" if ch["temporal"] else " (independent)"))
    return lines


def camera_ids(rig_json_file):
    rig = json.load(open(rig_json_file))
    return [c["id"] for c in rig["cameras"]]


def run_rank(args, rank, plan):
    """Renders this rank's chains: one context and one host thread per chain (the GPU overlaps the chains; PNG decode / encode of a chain
    overlaps the rendering of the others)."""
    import cv2
    import numpy as np
    from . import api
    root = args.root_dir
    eqr_dir, cube_dir = os.path.join(root, "eqr_frames"), os.path.join(root, "cube_frames")
    os.makedirs(eqr_dir, exist_ok=True)
    os.makedirs(cube_dir, exist_ok=True)
    if args.save_debug_images:
        print("[batch_process_video] --save_debug_images is ignored by this driver", file=sys.stderr)
    ids = camera_ids(args.rig_json_file)
    kw = quality_params(args.quality)
    want_cube = args.cubemap_width > 0 and args.cubemap_height > 0
    params_kw = dict(kw, enable_top=int(args.enable_top), enable_bottom=int(args.enable_bottom), enable_pole_removal=int(args.enable_pole_removal),
                     side_flow_alg=args.flow_alg, polar_flow_alg=args.flow_alg, poleremoval_flow_alg=args.flow_alg, want_cubemap=int(want_cube),
                     cubemap_photo=int(args.cubemap_format == "photo"))
    if want_cube:
        params_kw.update(cubemap_width=args.cubemap_width, cubemap_height=args.cubemap_height)
    device = int(os.environ.get("LOCAL_RANK", rank))
    chains = plan[rank]
    errors, done = [], [0]
    lock = threading.Lock()

    def load_frame(frame):
        ims = []
        for cid in ids:
            path = os.path.join(root, "rgb", cid, f"{frame:06d}.png")
            im = cv2.imread(path, cv2.IMREAD_COLOR)
            if im is None:
                raise RuntimeError("failed to load image: " + path)
            ims.append(im)
        return ims

    def work(chain):
        try:
            ctx = api.Context(device)
            ctx.load_rig(args.rig_json_file)
            if chain["temporal"]:
                ctx.set_temporal(True)
            if len(chains) <= 4:
                ctx.set_latency_mode(True)   # few frames in flight on this GPU: the latency variant of the sweeps wins; same bytes
            if args.enable_pole_removal:
                bottom, bottom2 = ctx.special_camera(1), ctx.special_camera(2)
                masks = [cv2.imread(os.path.join(root, "pole_masks", ctx.camera_id(c) + ".png"), cv2.IMREAD_COLOR) for c in (bottom, bottom2)]
                if masks[0] is None or masks[1] is None:
                    raise RuntimeError("missing or bad pole mask in " + os.path.join(root, "pole_masks"))
                ctx.set_pole_masks(*masks)
            params = api.default_params(**params_kw)
            for frame in chain["frames"]:
                out = ctx.renderStereoPanorama(params, load_frame(frame))
                eq, cube = out if isinstance(out, tuple) else (out, None)
                if not cv2.imwrite(os.path.join(eqr_dir, f"eqr_{frame:06d}.png"), eq):
                    raise RuntimeError("failed to write image: eqr_%06d.png" % frame)
                if cube is not None and not cv2.imwrite(os.path.join(cube_dir, f"cube_{frame:06d}.png"), cube):
                    raise RuntimeError("failed to write image: cube_%06d.png" % frame)
                with lock:
                    done[0] += 1
                    if args.verbose:
                        print(f"----------- [Render] rank {rank}: frame {frame} done ({done[0]} of {sum(len(c['frames']) for c in chains)})", flush=True)
            ctx.close()
        except Exception as e:   # reported by the main thread
            with lock:
                errors.append(f"chain starting at frame {chain['frames'][0]}: {e}")

    t0 = time.time()
    threads = [threading.Thread(target=work, args=(ch,)) for ch in chains]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise SystemExit("[batch_process_video] rank %d: %s" % (rank, "; ".join(errors)))
    n = sum(len(c["frames"]) for c in chains)
    if args.verbose and n:
        dt = time.time() - t0
        print(f"Render total runtime: {dt:.1f} sec (rank {rank}, {n} frames)\nAverage runtime: {dt / n:.3f} sec/frame", flush=True)


def main(argv=None):
    args = parse_args(argv)
    plan = make_plan(args)
    if args.dryrun:
        print("\n".join(describe(plan)))
        return 0
    if "RANK" in os.environ:   # launched by torchrun (or by the parent below): render this rank's share
        run_rank(args, int(os.environ["RANK"]), plan)
        return 0
    if args.gpus == 1:
        run_rank(args, 0, plan)
        return 0
    procs = []
    for r in range(args.gpus):   # one process per GPU; no collective is needed: every rank reads the rig JSON itself
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(args.gpus))
        procs.append(subprocess.Popen([sys.executable, "-m", "surround360_b200.batch_process_video"] + (argv if argv is not None else sys.argv[1:]), env=env))
    rc = 0
    for p in procs:
        rc = p.wait() or rc
    return rc


if __name__ == "__main__":
    sys.exit(main())
