"""Group the SASS of a kernel in an .ncu-rep into runs with similar execution counts (loop nests) and
print where the warp-instructions go:  python scripts/ncu_hot.py file.ncu-rep [kernel-substring] [--sass N]"""
import csv
import io
import subprocess
import sys


def main(path, pattern='', show=0):
    args = ['ncu', '-i', path, '--page', 'source', '--csv']
    out = subprocess.run(args, capture_output=True, text=True).stdout
    blocks = out.split('"Kernel Name",')
    for block in blocks[1:]:
        lines = block.splitlines()
        name = lines[0]
        if pattern not in name:
            continue
        rows = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
        header = rows[0]
        ia, isrc, iex, ism = header.index('Address'), header.index('Source'), header.index('Instructions Executed'), header.index('# Samples')
        ith = header.index('Avg. Threads Executed')
        recs = [(r[isrc].strip(), int(r[iex] or 0), int(r[ism] or 0), float(r[ith] or 0)) for r in rows[1:] if len(r) > iex]
        total = sum(r[1] for r in recs)
        samples = sum(r[2] for r in recs)
        print('====', name[:90], 'instructions', total, 'samples', samples)
        runs = []
        for idx, (src, ex, sm, th) in enumerate(recs):
            if runs and ex > 0 and abs(ex - runs[-1]['ex'])/max(ex, runs[-1]['ex'], 1) < 0.15:
                r = runs[-1]
                r['n'] += 1; r['tot'] += ex; r['sm'] += sm; r['end'] = idx; r['th'] += th*ex
            else:
                runs.append(dict(ex=ex, n=1, tot=ex, sm=sm, start=idx, end=idx, th=th*ex))
        for r in runs:
            if r['tot'] > 0.01*total:
                print('  sass[%4d..%4d] n=%4d  exec/instr %10d  share %5.1f%%  samples %5.1f%%  lanes %4.1f' % (
                    r['start'], r['end'], r['n'], r['ex'], 100*r['tot']/total, 100*r['sm']/max(samples, 1), r['th']/max(r['tot'], 1)))
                if show:
                    for k in range(r['start'], min(r['end'] + 1, r['start'] + show)):
                        print('        ', recs[k][0][:90], recs[k][1], recs[k][2])
        return


if __name__ == '__main__':
    show = 0
    argv = sys.argv[1:]
    if '--sass' in argv:
        k = argv.index('--sass')
        show = int(argv[k+1])
        argv = argv[:k] + argv[k+2:]
    main(argv[0], argv[1] if len(argv) > 1 else '', show)
