"""The reference's example flow (examples/implicit-recsys/{bpr,wmf,relmf}_example.py) on a synthetic MovieLens-shaped
dataset, B200 required:   python examples/quickstart.py --model bpr --max_epochs 50"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cymf_b200 as cymf  # noqa: E402

parser = argparse.ArgumentParser(description="")
parser.add_argument("--model", choices=["bpr", "wmf", "relmf"], default="bpr")
parser.add_argument("--dataset", default="ml-100k")
parser.add_argument("--max_epochs", type=int, default=300)
parser.add_argument("--num_components", type=int, default=20)
parser.add_argument("--learning_rate", type=float, default=1e-2)
parser.add_argument("--weight_decay", type=float, default=1e-2)
parser.add_argument("--num_threads", type=int, default=8)
args = parser.parse_args()

dataset = cymf.dataset.SyntheticMovieLens(args.dataset)
valid_evaluator = cymf.evaluator.AverageOverAllEvaluator(dataset.valid, dataset.train, metrics=["DCG"], k=5)
test_evaluator = cymf.evaluator.AverageOverAllEvaluator(dataset.test, dataset.train, k=5)
if args.model == "bpr":
    model = cymf.BPR(num_components=args.num_components, learning_rate=args.learning_rate, weight_decay=args.weight_decay)
elif args.model == "wmf":
    model = cymf.WMF(num_components=args.num_components, weight_decay=args.weight_decay)
else:
    model = cymf.RelMF(num_components=args.num_components, learning_rate=args.learning_rate,
                       weight_decay=args.weight_decay)
model.fit(dataset.train, num_epochs=args.max_epochs, num_threads=args.num_threads, valid_evaluator=valid_evaluator,
          early_stopping=True)
print(test_evaluator.evaluate(model.W, model.H))
