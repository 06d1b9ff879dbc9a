"""`python -m gprmax_b200 model.in -gpu [ids ...]` = `python -m gprMax model.in -gpu [ids ...]` on the B200 core."""
from .dropin import main

if __name__ == '__main__':
    main()
