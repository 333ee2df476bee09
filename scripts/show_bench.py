import json, sys
d = json.load(open(sys.argv[1]))
print('value %.4g e2e %.4g frac %.3f kernel_ms %.3f' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms']))
for k, v in d['extra'].items():
    if 'configs_per_s' in v:
        print(k, '%.4g cfg/s %.4g edges/s (uniform) %.4g edges/s (local)' % (v['configs_per_s'], v['edges_per_s'], v.get('local_edges_per_s', float('nan'))))
if 'knn_box_stacking_100k' in d['extra']:
    k = d['extra']['knn_box_stacking_100k']
    print('knn tensor %.1f ms exact %.1f ms' % (k['tensor_ms'], k['exact_ms']))
