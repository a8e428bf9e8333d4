"""The reference launcher's imports (pioneer/launch/pioneer_knm_train.py:10-11, pioneer/envs/bullet/__init__.py:1-2) resolve
against this repo when compat/ is on the path.  Run in a subprocess: the alias must not leak into this interpreter (the golden
generator imports the real reference package under the same name)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_paths_resolve():
    code = ("from pioneer.envs.pioneer import PioneerKinematicConfig\n"
            "from pioneer.envs.pioneer import PioneerKinematicEnv\n"
            "from pioneer.envs.bullet import BulletEnv, SimulationConfig, RenderConfig, Scene, World, Joint, Item, Pose, Velocity\n"
            "import pioneer_b200.envs.pioneer as p\n"
            "assert PioneerKinematicEnv is p.PioneerKinematicEnv and issubclass(PioneerKinematicEnv, BulletEnv)\n"
            "assert PioneerKinematicConfig().done_distance == 0.1 and SimulationConfig().frames_per_second == 24\n"
            "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "compat")]))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/")
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr[-2000:]
