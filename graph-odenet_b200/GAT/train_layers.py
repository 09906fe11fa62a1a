"""``python -m graph_odenet_b200.GAT.train_layers --dataset cora --runs 2`` -- GAT/train_layers.py (depth sweep + result pickles) on libgode."""
from ..train_layers import main as _main


def main(argv=None):
    return _main("GAT", argv)


if __name__ == "__main__":
    main()
