"""Mirror of the reference's process wrapper (lq_exec.py:5-71): LongQC runs the binaries through
``LqExec(bin).exec(*argv, out=, err=)`` and polls ``get_poll()``.  Kept byte-compatible in behaviour
so the drop-in can be exercised exactly the way longQC.py:438-445 does."""
import subprocess


class LqExec:
    def __init__(self, bin_path, logger=None):
        self.bin_path = bin_path
        self.logger = logger
        self.proc = None
        self._out = self._err = None

    def exec(self, *args, out=None, err=None):
        argv = [self.bin_path] + [str(a) for a in args]
        self._out = open(out, "w") if out else subprocess.DEVNULL
        self._err = open(err, "w") if err else subprocess.DEVNULL
        self.proc = subprocess.Popen(argv, stdout=self._out, stderr=self._err)
        return self.proc.pid

    def get_poll(self):
        rc = self.proc.poll()
        if rc is not None:
            for f in (self._out, self._err):
                if hasattr(f, "close"):
                    f.close()
        return rc

    def wait(self):
        rc = self.proc.wait()
        self.get_poll()
        return rc

    def get_pid(self):
        return self.proc.pid

    def get_bin_path(self):
        return self.bin_path
