"""``python -m graph_odenet_b200.GAT.train_res --model ode3 --dataset cora`` -- GAT/train_res.py on libgode."""
from ..train import main as _main


def main(argv=None):
    return _main("GAT", argv)


if __name__ == "__main__":
    main()
