"""Import-order shim: the reference's entry points start with `import isaacgym` (train.py:1, play.py:1) because Isaac
Gym must be imported before torch.  There is no Isaac Gym behind this implementation - the physics is
libb200t1.so - so this package only has to exist.  (The richer stub under oracle/shims/ is test infrastructure for
importing the REFERENCE's envs; it is never on the product's path.)"""
