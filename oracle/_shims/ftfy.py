def fix_text(s):
    return s
