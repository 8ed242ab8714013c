"""Parses the reference's recorded model.summary() (notebooks/Train/Train_tests.ipynb cell 9
output) into tests/golden/unet_summary.json -- the only known-answer the reference ships for the
U-Net half: layer order/types, per-layer param counts and the totals."""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
nb = json.load(open('/root/reference/notebooks/Train/Train_tests.ipynb'))
text = ''
for o in nb['cells'][9].get('outputs', []):
    if 'text' in o:
        text += ''.join(o['text'])
layers = []
for line in text.splitlines():
    m = re.match(r'^(\S+) \((\S+?)\)?\s+[\[(].*?\)\s+(\d+)\s', line)
    if m:
        layers.append({'name': m.group(1), 'type': m.group(2), 'params': int(m.group(3))})
tot = {k: int(re.search(k + r': ([\d,]+)', text).group(1).replace(',', ''))
       for k in ('Total params', 'Trainable params', 'Non-trainable params')}
json.dump({'layers': layers, 'totals': tot, 'config': {'DIM': [128, 128], 'DEPTH': 4, 'FILTERS': 32,
           'IMG_CHANNELS': 1, 'MASK_CLASSES': 2}}, open(os.path.join(HERE, 'unet_summary.json'), 'w'), indent=1)
print(len(layers), tot)
