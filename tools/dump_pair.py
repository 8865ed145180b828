"""Write a synthetic scan pair as raw float32 xyzw files for examples/lidar_odometry_sequence.c."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rolo_b200 import synth
sensor = sys.argv[1] if len(sys.argv) > 1 else "small"
out = sys.argv[2] if len(sys.argv) > 2 else "."
t, s, _ = synth.scan_pair(sensor)
t.tofile(os.path.join(out, "target.bin")); s.tofile(os.path.join(out, "source.bin"))
print(t.shape, s.shape)
