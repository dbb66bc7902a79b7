"""stub: utils/runner.py imports imageio at module level (video writing is out of scope)"""


def get_writer(*a, **k):
    raise RuntimeError("imageio stub: video recording is not available")
