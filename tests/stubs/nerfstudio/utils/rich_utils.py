class _Console:
    def log(self, *a, **k):
        pass

    print = log
    rule = log

    def input(self, *a, **k):
        return ""


CONSOLE = _Console()
