"""Extracts the numbers that the reference's own executed notebooks print for the hot path
 into tests/golden/reference_notebook_values.json.

Run in the build container, where the reference is mounted read-only:
    python tests/golden/extract_reference_goldens.py [/root/reference]
The JSON is committed; tests read only the JSON (the reference does not exist on the GPU box).
Nothing but printed output values is taken from the reference."""
import json
import os
import re
import sys

WANT = [
    # key, notebook, substring of the cell source, what the printed value is
    ("caches_dot_ones_surface", "caches.ipynb", "dot(os,os,cache)", "sum of ds on the circle R=1, ds target 1.4dx, dx=0.01"),
    ("caches_integrate_ones_surface", "caches.ipynb", "integrate(os,cache)", "same by integrate()"),
    ("caches_integrate_ones_grid", "caches.ipynb", "og = ones_grid(cache)", "interior area of the 406^2 grid, dx=0.01"),
    ("caches_integrate_ones_gridgrad", "caches.ipynb", "ovg = ones_gridgrad(cache)", "same per component"),
    ("caches_normals_head", "caches.ipynb", "normals(cache)", "first printed normals components (nx_k)"),
    ("caches_grid", "caches.ipynb", "g = PhysicalGrid(", "grid printout: size, origin index, dx, limits"),
    ("caches_circle", "caches.ipynb", "body = Circle(RadC", "number of points of Circle(1.0, 1.4dx)"),
    ("layers_dot_x_Rf_n", "Layers.ipynb", "dot(qx,Rf*nrm,g)", "integral of x n_x over the regularized circle ~ pi"),
    ("multbodies_circle", "multbodies.ipynb", "body = Circle(RadC", "number of points of Circle(0.5, 1.4dx)"),
    # time marching of the constrained heat equation (IF-HERK through ConstrainedSystems): temperature at the grid node
    # (-0.9, 0) after 51 and 54 steps of dt = 1e-4 on the 408^2 grid (dx = 0.01), circle R = 1, T+ = 0, T- = 1
    ("heatconduction_dt", "heatconduction.ipynb", "timestep_fourier(u0,sys)", "time step Fo dx^2 / kappa"),
    ("heatconduction_state_size", "heatconduction.ipynb", "state(u0)", "Nodes{Primal,408,408}: grid size"),
    ("heatconduction_T_t0051", "heatconduction.ipynb", "Tfcn(-0.9,0)", "T(-0.9, 0) at t = 0.0051"),
    ("heatconduction_T_t0054", "heatconduction.ipynb", "Tfcn_array[4](-0.9,0)", "T(-0.9, 0) at t = 0.0054"),
    ("heatconduction_maxvelocity", "heatconduction.ipynb", "maxvelocity(body,x,m)", "max surface speed of the deforming circle, its point index"),
    ("neumann_added_mass", "neumann.ipynb", "M = -integrate(df", "added-mass integral of the Neumann solve"),
    ("multbodies_volume", "multbodies.ipynb", "V3 = integrate(pointwise_dot(pts,nrm),cache,3)", "area of body 3 by the divergence theorem"),
]
NUM = re.compile(r"[-+]?(?:\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+(?:[eE][-+]?\d+)?)")


def cell_output(nb, needle):
    for ci, c in enumerate(nb["cells"]):
        if c["cell_type"] != "code" or needle not in "".join(c["source"]):
            continue
        for o in c.get("outputs", []):
            txt = "".join(o.get("text", []) or o.get("data", {}).get("text/plain", []))
            if txt:
                return ci, "".join(c["source"]), txt
    raise KeyError(needle)


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out = {}
    for key, nbname, needle, what in WANT:
        nb = json.load(open(os.path.join(ref, "examples", nbname)))
        ci, src, txt = cell_output(nb, needle)
        out[key] = {
            "notebook": f"examples/{nbname}", "cell": ci, "what": what, "source": src, "text": txt,
            # every number in the printout, and separately the lines that are nothing but one number
            "numbers": [float(x) for x in NUM.findall(txt)],
            "values": [float(l) for l in (ln.strip() for ln in txt.split("\n")) if NUM.fullmatch(l)],
        }
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_notebook_values.json")
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst)
    for k, v in out.items():
        print(k, v["values"][:6], v["numbers"][:6])


if __name__ == "__main__":
    main()
