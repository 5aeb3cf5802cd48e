class Error(Exception):
    pass


class NameNotFound(Error):
    pass
