"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

CPU oracle for the TUCH self-contact hot path.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; tuch_b200/ never does.

* oracle.clib     -- ctypes binding of the C restatement (contact_oracle.c)
* oracle.lbs      -- torch-CPU restatement of smplx==0.1.13 lbs + tuch/models/smpl.py wrapper
* oracle.losses   -- torch-CPU restatement of tuch/smplify/{losses,prior}.py, tuch/utils/geometry.py
* oracle.segments -- tuch/utils/segmentation.py
* oracle.smplify  -- tuch/smplify/smplifydc.py loop
* oracle.regressor-- tuch/train/loss.py contact_loss (+HD path), train_module.contact_from_verts

Parity pins: everything that exists in /root/reference is checked against outputs of the
reference's own Python recorded in tests/golden/ (tests/golden/make_golden.py).  The LBS
arithmetic lives in the un-vendored third-party smplx==0.1.13 (requirements.txt:13), absent
from /root/reference: for oracle.lbs parity is UNPINNED (restated from the published algorithm).
"""
