"""Drop-in import shim: `from model.TSNet import TSNet` / `from model.TSNet_pose import TSNet` -- the import
lines of the reference's quick_start1.py:3, demo/demo_face.py:17, demo/demo_pose.py:14 -- resolve to the B200 classes
when the repo root is on sys.path."""
