"""ExternalFileOracle (reference: util/oracle/ExternalFileOracle.py) is importable so that configs naming it load,
but fundamentals read from files are not on the B200 path: lowering accepts the SparseMeanRevertingOracle only."""


class ExternalFileOracle:
    def __init__(self, symbols):
        raise NotImplementedError("ExternalFileOracle is not supported by the B200 engine "
                                  "(use util.oracle.SparseMeanRevertingOracle)")
