"""refshim — TEST INFRASTRUCTURE, in-container only.

Lets the reference's own Python (`/root/reference/metadrive`, MetaDrive v0.4.2.2) import and run
without panda3d / gymnasium / shapely / pygame, so that its managers, IDM policy, navigation,
`perceive`, reward and observation code execute UNMODIFIED and produce golden vectors for
`tests/golden/` (plan items 1-3).

Three layers:
  * `stubs`   - a `sys.meta_path` finder that fabricates inert modules for rendering-only packages;
  * `pcore`   - a small semantic stand-in for the parts of `panda3d.core` the step path relies on
                (NodePath transforms, Vec3, BitMask32, TransformState, python tags);
  * `pbullet` - a recording + analytic stand-in for `panda3d.bullet` (shapes, bodies, BulletWorld
                queries, BulletVehicle), whose arithmetic lives in `physics.py`
                (an independent float64 restatement of the Bullet calls on the path).

Nothing in the product (`metadrive_ped_b200/`) imports this package.  It cannot run on the GPU box
(`/root/reference` does not exist there); only the fixtures it generated travel.
"""
import sys

REFERENCE_ROOT = "/root/reference"
_installed = False


def install():
    """Install stubs + fakes and put the reference on sys.path. Idempotent."""
    global _installed
    if _installed:
        return
    from . import stubs
    stubs.install()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True
    from . import patches
    patches.apply()
