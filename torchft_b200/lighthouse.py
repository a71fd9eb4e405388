"""``python -m torchft_b200.lighthouse --min_replicas N [--bind [::]:29510] ...``

Console entry for the Lighthouse quorum server (the reference installs the same thing as the
``torchft_lighthouse`` script, pyproject.toml:43-44). A Python-free binary is built at
``bin/torchft_b200_lighthouse``.
"""

from torchft_b200._C import lighthouse_main


def main() -> None:
    raise SystemExit(lighthouse_main())


if __name__ == "__main__":
    main()
