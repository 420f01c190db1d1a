"""Writes the synthetic benchmark scenes of BASELINE.md (C1-C5) in the reference's JSON schema
(SURVEY.md App. C).  Geometry follows the reference's shipped scenes with the mesh bodies removed."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def scene(name, method, domain_end, start, end, velocity, dt, translation=(0, 0, 0), visc_method="standard",
          viscosity=10.0, viscosity_b=5.0, g_upper=None, total_time=2.0):
    cfg = {"domainStart": [0.0, 0.0, 0.0], "domainEnd": list(domain_end), "addDomainBox": True, "particleRadius": 0.01,
           "fps": 60.0, "totalTime": total_time, "density0": 1000, "gravitation": [0.0, -9.81, 0.0],
           "simulationMethod": method, "viscosityMethod": visc_method, "timeStepSize": dt, "viscosity": viscosity,
           "viscosity_b": viscosity_b, "exportFrame": False, "exportPly": True, "exportObj": False}
    if g_upper is not None:
        cfg["gravitationUpper"] = g_upper
    block = {"objectId": 0, "start": list(start), "end": list(end), "translation": list(translation), "scale": [1, 1, 1],
             "velocity": list(velocity), "density": 1000.0, "color": [50, 100, 200], "entryTime": -1.0}
    with open(os.path.join(HERE, "scenes", name + ".json"), "w") as f:
        json.dump({"Configuration": cfg, "FluidBlocks": [block]}, f, indent=2)


if __name__ == "__main__":
    scene("dam_break_8k_wcsph", "wcsph", (1, 1, 1), (0.1, 0.1, 0.1), (0.495, 0.495, 0.495), (0, 0, 0), 4e-4)            # C1
    scene("dam_break_1m_wcsph", "wcsph", (8.5, 8.0, 2.0), (0.09, 0.2, 0.2), (1.7, 4.0, 1.8), (0, -0.5, 0), 4e-4, viscosity_b=0.3)   # C2
    scene("dam_break_1m_dfsph", "dfsph", (8.5, 8.0, 2.0), (0.09, 0.2, 0.2), (1.7, 4.0, 1.8), (0, -0.5, 0), 6e-4, viscosity_b=0.3)   # C2'
    scene("bath_500k_dfsph", "dfsph", (5, 3, 2), (0.3, 0.2, 0.5), (1.2, 2.8, 1.6), (0, -1, 0), 2e-3, translation=(0.2, 0, 0.2))   # C3
    scene("buckling_pcisph_implicit", "pcisph", (4, 20, 8), (1.12, 1, 1), (1.88, 12.2, 1.08), (0, -2.2, 0.75), 1e-3,
          visc_method="implicit", viscosity=1800.0, viscosity_b=1800.0, g_upper=2.5)                                           # C4
    scene("dam_break_10m_dfsph", "dfsph", (6, 6, 8), (0.2, 0.2, 0.2), (4.195, 5.195, 4.195), (0, 0, 0), 6e-4)                    # C5
