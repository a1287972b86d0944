"""Print the headline numbers of the bench lines in a gpurun_out log."""
import json
import sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l)
        la = d['chain']['lookahead']
        n = la['window_passes']
        print("value %.4f e2e %.4f steps %.3f refit %.3f pass %.0f us frac %.3f launches %d" % (
            d['value'], d['e2e']['value'], d['chain']['steps_s_per_sweep'], d['chain']['refit_s_per_sweep'],
            d['roofline']['avg_launch_us'], d['roofline']['frac'], d['gpu_launches']))
        print("   per batch us: open %.0f chain %.0f ensure %.0f correct %.0f commit %.0f | rescoring passes %d replayed %d chained %d clusters %d" % (
            la['open_s'] / n * 1e6, la.get('chained_wait_s', 0) / n * 1e6, la['rescore_s'] / n * 1e6, la['correct_s'] / n * 1e6,
            la['commit_s'] / n * 1e6, la['rescoring_passes'], la['replayed'], la.get('chained_corrections', 0),
            d['chain']['active_clusters']))
    else:
        print(l.rstrip()[:300])
