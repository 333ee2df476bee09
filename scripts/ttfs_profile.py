"""cProfile of one batch-PRM run on the B200 backend: python scripts/ttfs_profile.py [scene] [n0] [t0]"""
import cProfile, pstats, sys
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
import ttfs
scene = sys.argv[1] if len(sys.argv) > 1 else "box_stacking"
n0 = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
t0 = int(sys.argv[3]) if len(sys.argv) > 3 else 600
ttfs.run(scene, "b200", 99, 200, 30, 30)
pr = cProfile.Profile()
pr.enable()
r = ttfs.run(scene, "b200", 0, n0, t0, 120)
pr.disable()
print(r)
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
