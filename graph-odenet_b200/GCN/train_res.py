"""``python -m graph_odenet_b200.GCN.train_res --model ode3 --dataset cora`` -- GCN/train_res.py on libgode."""
from ..train import main as _main


def main(argv=None):
    return _main("GCN", argv)


if __name__ == "__main__":
    main()
