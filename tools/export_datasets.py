"""Write the synthetic benchmark line sets as .binlines + datasets.json + a replay script, so that an unmodified LineVis on a
Vulkan machine can render exactly these inputs (closes the 'parity unpinned' gap from the reference side)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linevis_b200 import scenes, binlines

ap = argparse.ArgumentParser()
ap.add_argument("--out", default="exported_datasets")
ap.add_argument("--configs", nargs="+", default=["config2"])
args = ap.parse_args()
os.makedirs(os.path.join(args.out, "flow"), exist_ok=True)
gens = {"config2": ("B200 Helix 100k", scenes.helix_lines, {}), "config3": ("B200 Random 1M", scenes.random_segments, {}),
        "config5": ("B200 Curl Noise 10M", scenes.curl_noise_streamlines, {})}
entries = []
for c in args.configs:
    name, fn, kw = gens[c]
    pos, attr, seg = fn(**kw)
    path = os.path.join("flow", c + ".binlines")
    binlines.write_binlines(os.path.join(args.out, path), binlines.polylines_from_segments(pos, attr, seg), vertices_normalized=True)
    entries.append(dict(name=name, filename=path, linewidth=scenes.LINE_WIDTH, attributes="Attribute"))
open(os.path.join(args.out, "datasets.json"), "w").write(binlines.datasets_json(entries))
open(os.path.join(args.out, "b200_parity.py"), "w").write('''import g

def init_scene():
    g.set_duration(0)
    g.set_dataset('%s')
    g.set_renderer('Vulkan Ray Tracer')
    g.set_rendering_algorithm_settings({
        'line_width': 0.002, 'depth_cue_strength': 0.0,
        'ambient_occlusion_mode': 'RTAO', 'ambient_occlusion_strength': 1.0, 'ambient_occlusion_gamma': 1.0,
        'ambient_occlusion_iterations': 1, 'ambient_occlusion_samples_per_frame': 64, 'ambient_occlusion_radius': 0.1,
        'ambient_occlusion_distance_based': True, 'use_jittered_primary_rays': True,
        'use_analytic_intersections': True, 'num_samples_per_frame': 1, 'num_accumulated_frames': 1,
    })
    g.set_transfer_function('Standard.xml')
    g.set_duration(6)
''' % entries[-1]["name"])
print("wrote", args.out, [e["filename"] for e in entries])
