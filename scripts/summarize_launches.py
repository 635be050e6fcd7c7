"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path, top=30):
    with open(path) as handle:
        lines = [line for line in handle if not line.startswith('==')]
    total = collections.defaultdict(float)
    count = collections.Counter()
    big = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        value = float(row['Metric Value'].replace(',', ''))
        unit = row['Metric Unit']
        value = value/1000 if unit == 'ns' else value*1000 if unit == 'ms' else value
        key = row['Kernel Name'].split('(')[0][:80]
        total[key] += value
        count[key] += 1
        big[key].append(value)
    everything = sum(total.values())
    print('total %.1f us over %d launches' % (everything, sum(count.values())))
    for key, value in sorted(total.items(), key=lambda kv: -kv[1])[:top]:
        durations = sorted(big[key])
        print('%10.1f us %5.1f%%  n=%4d  avg %8.2f  max %8.2f  median %8.2f  %s' % (
            value, 100*value/everything, count[key], value/count[key], durations[-1], durations[len(durations)//2], key))


if __name__ == '__main__':
    main(sys.argv[1])
