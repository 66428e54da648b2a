#!/usr/bin/env perl
# perl/bin/ppp_hmmer_b200.pl -- command-line drop-in for ProPhylo's bin/ppp_hmmer.pl on the B200 path.
#
# One query profile against the hits of one hmmsearch 3 report: same options and stdout as the reference
# (/root/reference/bin/ppp_hmmer.pl:163-251).  The hit list above the cut-off (full-sequence scores, or best-domain scores
# with --domain) becomes the target -- distinct taxa in first-hit order, the HASH branch of ppp.pm:147-149 -- and the pair is
# scored on the GPU through PhyloProf::Algorithm::ppp_b200 (XS -> libppp_b200.so) instead of the Perl sweep (:220-222).
# Parsing and gi -> taxon lookups (SeqToolBox::Taxonomy) stay on the host.  There is no CPU fallback.
#
#   perl ppp_hmmer_b200.pl -p <profile> --hmm <hmmsearch output> [-l rank] [-t trusted cutoff] [--prob P] [--domain]
use strict;
use warnings;
use Getopt::Long;
use FindBin;
use lib "$FindBin::Bin/../lib";
use PhyloProf::B200;
use PhyloProf::B200::Host qw(parse_hmmsearch3 above_cutoff);
use PhyloProf::Algorithm::ppp_b200;

my ( $help, $profile, $hmmer_result, $level, $trusted_cutoff, $probability, $domain ) = ( 0, '', '', '', 0, 0, 0 );
GetOptions( 'help|h' => \$help, 'profile|p=s' => \$profile, 'hmmer_result|hmm=s' => \$hmmer_result, 'level|l=s' => \$level,
            'trusted_cutoff|t=s' => \$trusted_cutoff, 'prob=s' => \$probability, 'domain' => \$domain )
    or die "usage: ppp_hmmer_b200.pl -p <profile> --hmm <hmmsearch output> [-l rank] [-t cutoff] [--prob P] [--domain]\n";
die "usage: ppp_hmmer_b200.pl -p <profile> --hmm <hmmsearch output> [-l rank] [-t cutoff] [--prob P] [--domain]\n" if $help;
die "HMMER file not give" unless $hmmer_result && -s $hmmer_result;

eval { require SeqToolBox::Taxonomy; 1 } or die "ppp_hmmer_b200.pl: needs SeqToolBox::Taxonomy (the reference's lib/) on \@INC: $@";
eval { require PhyloProf::Profile;   1 } or die "ppp_hmmer_b200.pl: needs PhyloProf::Profile (the reference's lib/) on \@INC: $@";
my $taxonomy = SeqToolBox::Taxonomy->new();
my @pargs = ( -file => $profile, -taxonomy => $taxonomy );
push @pargs, ( -rank => $level ) if $level;
my $ref_profile = PhyloProf::Profile->new(@pargs);

my $report = parse_hmmsearch3($hmmer_result);
my $scores = $domain ? $report->{domain} : $report->{full};    # :246-296
my @hits   = @{ above_cutoff( $scores, $trusted_cutoff ) };
my ( %rank, $yes );                                            # HMMERResult.pm:162-242: taxon => rank of its first hit
for my $i ( 0 .. $#hits ) {
    my $t = $taxonomy->get_taxon( $hits[$i] );
    next unless $t;
    next if exists $rank{$t};
    $rank{$t} = $i + 1;
    $yes++;
}
my $target = PhyloProf::Profile->new( -raw => \%rank, -yes => $yes || 0, -no => 0 );
my @param = ( -prof1 => $ref_profile, -prof2 => $target );
push @param, ( -prob => $probability ) if $probability > 0;
push @param, ( -rank => $level ) if $level;
my $alg = PhyloProf::Algorithm::ppp_b200->new(@param);
$alg->{_dbh} = $taxonomy;
my ( $score, $nyes, $number, $depth, $p ) = $alg->get_match_score();
print "Score\t#yes\t#Toal\t#Depth\t#Prob\n";
die "Can't take log of 0\n" if $score == 0;
printf "%s\t%s\t%s\t%s\t%s\n", sprintf( "%.3f", log($score) / log(10) ), $nyes, $number, $depth, sprintf( "%.2f", $p );

# the first $depth hits of the report (the reference indexes the raw hit list with a depth that counts distinct taxa, :228-243)
my @lines;
for my $i ( 0 .. $depth - 1 ) {
    my $hit   = $hits[$i];
    my $taxon = defined $hit ? $taxonomy->get_taxon($hit) : undef;
    my $pos   = ( defined $taxon && $ref_profile->is_yes($taxon) ) ? 'Y' : 'N';
    no warnings 'uninitialized';
    push @lines, join( "\t", $taxon, $hit, $pos, $scores->{$hit}, $report->{desc}{$hit} );
}
print "Taxon\tGI\tPositive\tHMMER_score\tDesc\n";
print join( "\n", @lines );
print "\n";
exit(0);
