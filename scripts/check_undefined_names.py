"""Tiny pyflakes stand-in (no linter is installed in the image): reports names that are loaded but bound nowhere in the enclosing scopes.
usage: python scripts/check_undefined_names.py paddle_b200   (class-scope names used as defaults / decorators show up as false positives)"""
import ast, glob, builtins, sys
B = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__path__", "__spec__", "__package__", "__builtins__", "__class__"}
def bound_names(node):
    """names bound directly in this scope (not descending into nested function/class scopes)."""
    out = set()
    def visit(n, top=True):
        for c in ast.iter_child_nodes(n):
            if isinstance(c, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                out.add(c.name)
                for d in c.decorator_list: visit(d, False)
                continue
            if isinstance(c, ast.Lambda): continue
            if isinstance(c, (ast.Import, ast.ImportFrom)):
                for a in c.names: out.add((a.asname or a.name).split(".")[0])
            if isinstance(c, ast.Name) and isinstance(c.ctx, (ast.Store, ast.Del)): out.add(c.id)
            if isinstance(c, ast.ExceptHandler) and c.name: out.add(c.name)
            if isinstance(c, (ast.Global, ast.Nonlocal)): out.update(c.names)
            if isinstance(c, ast.arg): out.add(c.arg)
            if isinstance(c, (ast.MatchAs,)) and c.name: out.add(c.name)
            visit(c, False)
    visit(node)
    return out
def check(path):
    src = open(path).read()
    tree = ast.parse(src)
    if any(isinstance(n, ast.ImportFrom) and any(a.name == "*" for a in n.names) for n in ast.walk(tree)): star = True
    else: star = False
    probs = []
    def scope(node, env):
        names = bound_names(node)
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
            a = node.args
            for x in a.posonlyargs + a.args + a.kwonlyargs: names.add(x.arg)
            if a.vararg: names.add(a.vararg.arg)
            if a.kwarg: names.add(a.kwarg.arg)
        env2 = env | names if not isinstance(node, ast.ClassDef) else env   # class scope names are not visible in methods
        local_env = env | names
        def walk(n):
            for c in ast.iter_child_nodes(n):
                if isinstance(c, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda)):
                    for d in getattr(c, "decorator_list", []): walk_expr(d)
                    for dflt in c.args.defaults + [k for k in c.args.kw_defaults if k is not None]: walk_expr(dflt)
                    scope(c, env2)
                elif isinstance(c, ast.ClassDef):
                    for d in c.decorator_list + c.bases: walk_expr(d)
                    scope(c, env2)
                elif isinstance(c, (ast.ListComp, ast.SetComp, ast.DictComp, ast.GeneratorExp)):
                    scope(c, local_env)
                else:
                    if isinstance(c, ast.Name) and isinstance(c.ctx, ast.Load) and c.id not in local_env and c.id not in B:
                        probs.append((c.lineno, c.id))
                    walk(c)
        def walk_expr(e):
            if isinstance(e, ast.Name) and isinstance(e.ctx, ast.Load) and e.id not in local_env and e.id not in B: probs.append((e.lineno, e.id))
            walk(e)
        walk(node)
    scope(tree, set())
    return [] if star else probs
tot = 0
for f in sorted(glob.glob(sys.argv[1] + "/**/*.py", recursive=True)):
    try: p = check(f)
    except Exception as e: print("ERR", f, e); continue
    for ln, name in p:
        print(f"{f}:{ln}: undefined name {name}"); tot += 1
print("total", tot)
