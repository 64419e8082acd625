def colorize(string, color=None, bold=False, highlight=False):
    return string
