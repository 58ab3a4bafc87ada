"""Plot helpers with the names the reference's scripts pick up through `from MPPI import *`
(ds_mppi/functions/plots.py).  Visualisation is out of scope for this package: with matplotlib installed
these draw a minimal 2-D view; without it they return inert handles so the scripts run headless."""
try:  # pragma: no cover - matplotlib is optional
    import matplotlib.pyplot as plt
    _HAVE_MPL = True
except Exception:  # noqa: BLE001
    _HAVE_MPL = False

    class _Inert:
        """Swallows every attribute access / call; `plt.plot(...)` unpacks to one handle."""
        def __getattr__(self, name):
            return self

        def __call__(self, *a, **k):
            return self

        def __iter__(self):
            return iter((self,))

        def __getitem__(self, i):
            return self

    plt = _Inert()


def _line(fig_id, style, **kw):
    if _HAVE_MPL:
        plt.ion()
        plt.figure(fig_id)
    (h,) = plt.plot([], [], style, **kw)
    return h


def _axes(fig_id, lims, labels, aspect='equal'):
    if not _HAVE_MPL:
        return
    plt.ion()
    ax = plt.figure(fig_id).add_subplot(111)
    ax.set_xlim(*lims[0]); ax.set_ylim(*lims[1])
    ax.set_xlabel(labels[0]); ax.set_ylabel(labels[1])
    ax.set_aspect(aspect)


def init_robot_plot(links, xmin, xmax, ymin, ymax):
    _axes(1, ((xmin, xmax), (ymin, ymax)), ('x, m', 'y, m'))
    return _line(1, 'o-', linewidth=3, markersize=5)


def init_jpos_plot(xmin, xmax, ymin, ymax):
    _axes(2, ((xmin, xmax), (ymin, ymax)), ('First Joint, radians', 'Second Joint, radians'))
    return _line(2, 'o', markersize=7)


def init_toy_plot(xmin, xmax, ymin, ymax):
    _axes(1, ((xmin, xmax), (ymin, ymax)), ('X', 'Y'))
    return _line(1, '*', markersize=5)


def init_kernel_means(n_kernel_max):
    return [_line(1, '-o', color='g', markersize=2, linewidth=1.5) for _ in range(n_kernel_max)]


def upd_jpos_plot(jpos, ln):
    ln.set_data(jpos[0], jpos[1])
    plt.draw()
    return 0


def upd_toy_h(coord, ln):
    ln.set_data(coord[0], coord[1])
    plt.draw()
    return 0


def upd_r_h(links, ln):
    xs = [links[0][0, 0]] + [link[-1, 0] for link in links]
    ys = [links[0][0, 1]] + [link[-1, 1] for link in links]
    ln.set_data(xs, ys)
    plt.draw()
    return 0


def plot_circ(c, r):
    if not _HAVE_MPL:
        return plt
    circ = plt.Circle(c[0:2], r, color='r', fill=False, linewidth=2)
    plt.figure(1).get_axes()[0].add_patch(circ)
    return circ


def plot_obs_init(obstacles):
    return [plot_circ(o[0:2], o[-1]) for o in obstacles]


def plot_obs_update(o_h, obstacles):
    for h, o in zip(o_h, obstacles):
        h.center = o[0:2]


def init_robot_plot3d(xmin, xmax, ymin, ymax, zmin, zmax, width, color, markersize):
    if not _HAVE_MPL:
        return plt
    plt.ion()
    ax = plt.figure(1).add_subplot(111, projection='3d')
    ax.set_xlim(xmin, xmax); ax.set_ylim(ymin, ymax); ax.set_zlim(zmin, zmax)
    (h,) = ax.plot3D([], [], [], 'o-', color=color, linewidth=width, markersize=markersize)
    return h


def init_line3d(width, color, markersize):
    if not _HAVE_MPL:
        return plt
    (h,) = plt.figure(1).get_axes()[0].plot3D([], [], [], 'o-', color=color, linewidth=width, markersize=markersize)
    return h


def upd_r_h3d(links, ln):
    pts = [[0, 0, 0], [float(v) for v in links[0][0]]] + [[float(v) for v in link[-1]] for link in links]
    xs, ys, zs = zip(*pts)
    ln.set_xdata(xs); ln.set_ydata(ys); ln.set_3d_properties(zs)
    plt.draw()
    return 0
