import os, sys, tempfile, pathlib, time, numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
os.environ['FDES_B200_TIMING'] = '1'
import fdes_b200 as fb
from fdes_b200 import specimens
tmp = pathlib.Path(tempfile.mkdtemp()); os.chdir(tmp)
cnf = tmp / 'si.cnf'
atoms = specimens.config_si001_1024(cnf, frozen_phonons=16)
a6 = np.ascontiguousarray(atoms, np.float32)
img = np.zeros((1, 512, 512), np.float32)
for i in range(5):
    t = time.perf_counter()
    fb.cuda_FDES(0, 0, str(cnf), str(tmp/'M.bin'), str(tmp/'r.emd'), a6, len(a6), img)
    print('call', i, (time.perf_counter()-t)*1e3, 'ms', img.mean())
