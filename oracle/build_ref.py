#!/usr/bin/env python
"""Recipe for oracle/_ref: the UNMODIFIED reference (ShangtongZhang/DistributedES, pure Python) placed where it can
travel to the GPU box.  TEST/BENCH INFRASTRUCTURE.

    python oracle/build_ref.py        # copies the five NES source files from /root/reference into oracle/_ref/

oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored, so the
files travel with the snapshot like the built .so.  __graft_entry__.build() runs this where /root/reference exists;
on the GPU box the copy made here is used as is.  Nothing under distributedes_b200/ reads it; bench.py's
`--impl reference` arm and `cpu_baseline` leg run it through oracle/ref_cpu_baseline.py.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
OUT = os.path.join(HERE, '_ref')
FILES = ['natural_es.py', 'utils.py', 'model.py', 'config.py', 'LICENSE']


def build(ref=REF, out=OUT):
    if not os.path.isdir(ref):
        return None
    os.makedirs(out, exist_ok=True)
    manifest = {}
    for f in FILES:
        src = os.path.join(ref, f)
        if not os.path.exists(src):
            continue
        shutil.copyfile(src, os.path.join(out, f))
        manifest[f] = hashlib.sha256(open(src, 'rb').read()).hexdigest()
    json.dump({'source': ref, 'sha256': manifest}, open(os.path.join(out, 'MANIFEST.json'), 'w'), indent=1)
    return out


if __name__ == '__main__':
    r = build()
    print(r if r else 'no %s here: oracle/_ref left as it is' % REF)
    sys.exit(0)
