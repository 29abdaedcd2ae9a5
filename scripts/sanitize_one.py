import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import athenapk_b200 as apk
nx = tuple(int(x) for x in os.environ.get("NX", "20,10,10").split(","))
mb = tuple(int(x) for x in os.environ.get("MB", "20,10,10").split(","))
s = apk.HydroSim(fp_mode="fast", xmin=(0.0, 0.0, 0.0), xmax=(1.0, 1.0, 1.0), gamma=5.0 / 3.0, cfl=0.3,
                 fluid="glmmhd", recon=os.environ.get("RECON", "wenoz"), riemann="hlle",
                 integrator=os.environ.get("INTEGRATOR", "rk1"), nx=nx, mb=mb, ng=3)
s.pgen("random", (3, 0.2)); s.initialize(); s.execute(None, 1); print("ok", s.dt); s.close()
