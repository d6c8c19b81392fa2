"""Generates rig JSON files in the reference's format (surround360_render/RIG_JSON.md):
a ring of RECTILINEAR side cameras plus FTHETA top / bottom / secondary-bottom cameras.
The default arguments reproduce the geometry of the reference's 17-camera rig
(res/config/camera_rig.json: 14 side cameras on a 21.8 cm ring, clockwise from +X, 2048^2,
focal 1269.58; fisheyes at z = +-13.1, focal 483.76, fov 1.61443 rad)."""
import json
import math


def make_rig(n_side=14, ring_radius=21.799999237060547, resolution=2048, side_focal=1269.580673376528, fisheye_focal=483.76220324,
             fisheye_fov=1.61443, pole_z=13.1, bottom2_y=9.8, side_distortion=(0.0, 0.0)):
    cams = []
    c = resolution / 2
    for i in range(n_side):
        a = -2 * math.pi * i / n_side  # clockwise seen from above
        fwd = [math.cos(a), math.sin(a), 0.0]
        right = [math.sin(a), -math.cos(a), 0.0]
        cams.append({"version": 1, "type": "RECTILINEAR", "group": "side camera", "id": f"cam{i + 1}",
                     "origin": [ring_radius * fwd[0], ring_radius * fwd[1], 0.0], "forward": fwd, "up": [0.0, 0.0, 1.0],
                     "right": right, "resolution": [resolution, resolution], "principal": [c, c],
                     "focal": [side_focal, -side_focal], "distortion": list(side_distortion)})

    def fisheye(cid, origin, fwd, up, right):
        return {"version": 1, "type": "FTHETA", "group": "", "id": cid, "origin": origin, "forward": fwd, "up": up,
                "right": right, "resolution": [resolution, resolution], "principal": [c, c],
                "focal": [fisheye_focal, -fisheye_focal], "distortion": [0, 0], "fov": fisheye_fov}

    cams.append(fisheye("cam0", [0, 0, pole_z], [0, 0, 1], [0, 1, 0], [-1, 0, 0]))
    cams.append(fisheye(f"cam{n_side + 1}", [0, 0, -pole_z], [0, 0, -1], [0, 1, 0], [1, 0, 0]))
    cams.append(fisheye(f"cam{n_side + 2}", [0, bottom2_y, -pole_z], [0, 0, -1], [0, -1, 0], [-1, 0, 0]))
    return {"cameras": cams}


def write_rig(path, **kw):
    with open(path, "w") as f:
        json.dump(make_rig(**kw), f, indent=1)


if __name__ == "__main__":
    import sys
    write_rig(sys.argv[1] if len(sys.argv) > 1 else "rig17.json")
