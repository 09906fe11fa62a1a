"""``python -m graph_odenet_b200.GCN.train_layers --dataset cora --runs 2`` -- GCN/train_layers.py (depth sweep + result pickles) on libgode."""
from ..train_layers import main as _main


def main(argv=None):
    return _main("GCN", argv)


if __name__ == "__main__":
    main()
