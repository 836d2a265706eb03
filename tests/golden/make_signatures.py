"""Records the call signatures of the reference's hot-path boundary by parsing its
sources with `ast` (nothing is imported or executed).  Run in the build container:
    python tests/golden/make_signatures.py   ->   tests/golden/signatures.json
tests/test_signatures_cpu.py holds the tuch_b200 mirrors to them (same names, order and defaults; the
mirrors may append keyword-only conveniences after the reference's parameters)."""
import ast
import json
import os

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
TARGETS = {
    'tuch/utils/contact.py': ['batch_pairwise_dist', 'solid_angles', 'winding_numbers'],
    'tuch/utils/segmentation.py': ['BodySegment.__init__', 'BodySegment.has_self_isect', 'BatchBodySegment.__init__',
                                   'BatchBodySegment.batch_has_self_isec'],
    'tuch/utils/geometry.py': ['batch_rodrigues', 'rot6d_to_rotmat', 'perspective_projection', 'estimate_translation'],
    'tuch/models/smpl.py': ['SMPL.forward'],
    'tuch/smplify/prior.py': ['MaxMixturePrior.__init__', 'MaxMixturePrior.forward'],
    'tuch/smplify/losses.py': ['gmof', 'contact_fitting_loss', 'camera_fitting_loss', 'angle_prior', 'body_fitting_loss'],
    'tuch/smplify/smplifydc.py': ['SMPLifyDC.__init__', 'SMPLifyDC.__call__', 'SMPLifyDC.get_fitting_loss'],
    'tuch/train/loss.py': ['batch_face_normals', 'RegressorLoss.__init__', 'RegressorLoss.forward', 'RegressorLoss.contact_loss'],
    'tuch/train/fits_dict.py': ['FitsDict.flip_pose', 'FitsDict.rotate_pose', 'FitsDict.__getitem__', 'FitsDict.__setitem__'],
    'tuch/train/train_module.py': ['TUCH.contact_from_verts', 'TUCH.__init__', 'TUCH.get_verts_in_contact',
                                   'TUCH.forward_train_step'],
    'tuch/eft/loss.py': ['EFTLoss.contact_loss'],
}


def signature(fn):
    a = fn.args
    names = [x.arg for x in a.posonlyargs + a.args]
    defaults = [None] * (len(names) - len(a.defaults)) + [ast.unparse(d) for d in a.defaults]
    return dict(params=names, defaults=defaults, vararg=a.vararg.arg if a.vararg else None,
                kwarg=a.kwarg.arg if a.kwarg else None, line=fn.lineno)


def main():
    out = {}
    for path, wanted in TARGETS.items():
        tree = ast.parse(open(os.path.join(REF, path)).read())
        found = {}
        for node in tree.body:
            if isinstance(node, ast.FunctionDef):
                found[node.name] = node
            elif isinstance(node, ast.ClassDef):
                for sub in node.body:
                    if isinstance(sub, ast.FunctionDef):
                        found['%s.%s' % (node.name, sub.name)] = sub
        for w in wanted:
            out['%s:%s' % (path, w)] = signature(found[w])
    with open(os.path.join(HERE, 'signatures.json'), 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(len(out), 'signatures')


if __name__ == '__main__':
    main()
