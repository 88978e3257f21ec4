"""placeholder filled in with the SCS fixtures"""


def main(only=""):
    return None
