import os, sys, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
shutil.copy(os.path.join(ROOT, "lifcal_b200", "liblfba_prof.so"), os.path.join(ROOT, "lifcal_b200", "liblfba.so"))
from lifcal_b200 import api, capi
for name in sys.argv[1:]:
    sc = capi.make_scene(int(name[-1]), order=1)
    ds = api.DeviceSolver(sc.problem, api.default_options(max_num_iterations=1))
    ds.set_parameters(sc.camera_init, sc.views_init, sc.points_init)
    print(name, flush=True)
    s = ds.run()
    ds.close()
