"""Generate the flattened-model fixtures that travel to the GPU box.

The URDF files live in the reference tree (/root/reference/models), which does not exist on the GPU
box; this script is run in the build container and its JSON outputs are committed next to it.
    python tests/golden/make_models.py
Known answers checked here (reference unittest/urdf.cpp:85,252; unittest/sample-models.cpp:43-44,70-71,92-93).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from pinocchio_b200.model import (JOINT_FREEFLYER, buildModelFromUrdf, buildSampleModelHumanoid,  # noqa: E402
                                  buildSampleModelHumanoidRandom, buildSampleModelManipulator)

REF = "/root/reference/models"
OUT = os.path.join(HERE, "models")


def main():
    os.makedirs(OUT, exist_ok=True)
    models = {}
    sh = buildModelFromUrdf(os.path.join(REF, "simple_humanoid.urdf"), JOINT_FREEFLYER)
    assert (sh.nq, sh.nv, sh.njoints) == (36, 35, 31), (sh.nq, sh.nv, sh.njoints)
    sh_noff = buildModelFromUrdf(os.path.join(REF, "simple_humanoid.urdf"))
    assert sh_noff.njoints == 30 and sh_noff.nq == 29
    models["simple_humanoid_ff"] = sh
    talos = buildModelFromUrdf(
        os.path.join(REF, "example-robot-data/robots/talos_data/robots/talos_reduced.urdf"), JOINT_FREEFLYER)
    assert (talos.nq, talos.nv, talos.njoints) == (39, 38, 34), (talos.nq, talos.nv, talos.njoints)
    models["talos_reduced_ff"] = talos
    man = buildSampleModelManipulator()
    assert (man.nq, man.nv) == (6, 6)
    models["manipulator"] = man
    hum = buildSampleModelHumanoid(True)
    assert (hum.nq, hum.nv) == (35, 34)
    models["humanoid"] = hum
    hr = buildSampleModelHumanoidRandom(True, seed=0)
    assert (hr.nq, hr.nv) == (33, 32)
    models["humanoid_random"] = hr
    for name, m in models.items():
        assert m.is_compact()
        with open(os.path.join(OUT, name + ".json"), "w") as fh:
            fh.write(m.to_json())
        print(name, m, "depth", max(m.depth()), "joints:", " ".join(m.names[1:8]), "...")


if __name__ == "__main__":
    main()
